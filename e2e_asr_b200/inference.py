"""Eval-mode decoding on the GPU: greedy (reference eval graph).
Beam search lives in beam_search.py."""
import torch

from . import ops
from ._lib import call


def greedy_decode_logits(v, decoder_inp, lens_i32, U, enc, enc_len_i32):
    """AttnDecoder eval path (attn_decoder.py:107,128-129; decoder.py:139-153):
    step 0 reads decoder_inp[0] (GO), later steps embed argmax(logits_{t-1}).
    Always runs U = max_output steps (seq2seq_model.py:191-193).  Returns logits
    [(U*B), V].  Same kernels as training, one step at a time."""
    return _stepwise_decode(v, decoder_inp, lens_i32, U, enc, enc_len_i32)[0]


def sample_decode_ids(v, decoder_inp, lens_i32, U, enc, enc_len_i32, samp_prob, seed, stream, lm_drop=None):
    """Scheduled sampling, first pass (attn_decoder.py:130-139, decoder.py:155-180): realise the decoder's input ids.
    Step 0 reads GO; at each step t >= 1 ONE scalar uniform u_t decides for the whole batch: the ground-truth
    decoder_inp[t] if u_t < 1 - samp_prob, else a multinomial draw from the previous step's logits.  The draw is not
    differentiable (the gradient reaches only the embedding row of the id that was fed), so the training step that
    follows is the teacher-forced step on these ids.  Returns ids [U, B] int64 on the device.

    Builder-defined randomness (TF's stateful generators cannot be reproduced): u_t = word 0 of
    philox(counter=(t, 200 + stream, 0, 0), key=seed) * 2^-32; row b of step t draws with
    counter (t*B + b, 300 + stream, 0, 0) through e2e_sample_rows.  Steps after the last sampled one need no logits
    and are not run."""
    from .host_utils import philox_uniform
    use_sample = [False] + [not (philox_uniform(t, 200 + stream, seed) < 1.0 - samp_prob) for t in range(1, U)]
    ids = decoder_inp[:U].clone()
    if not any(use_sample):
        return ids
    last = max(t for t in range(U) if use_sample[t])
    _stepwise_decode(v, decoder_inp, lens_i32, last, enc, enc_len_i32, lm_drop=lm_drop,
                     rule=dict(use_sample=use_sample, seed=seed, offset=300 + stream, ids=ids))
    return ids


def _stepwise_decode(v, decoder_inp, lens_i32, U, enc, enc_len_i32, lm_drop=None, rule=None):
    """One decoder step at a time with the training kernels.  rule=None: greedy.  Otherwise the scheduled-sampling
    rule above: after step t the input of step t+1 is written to rule["ids"][t+1]."""
    dev = enc.device
    f32 = dict(dtype=torch.float32, device=dev)
    st = ops._dev_state(dev)
    B, Tn, D = enc.shape
    V, E = v["emb"].shape
    Hl, Hd, A = v["lm_k"].shape[1] // 4, v["dec_k"].shape[1] // 4, v["q_k"].shape[1]
    enc_flat, r0, r1 = ops.flat_rows(enc)
    assert r1 == 1
    Tp = r0
    HF = ops.gemm(enc_flat, v["attn_w"].view(D, A))
    logits = torch.zeros((U * B, V), **f32)
    tok = decoder_inp[0].contiguous()
    u = torch.empty((B, E), **f32)
    xl = torch.zeros((B, E + Hl), **f32)      # (u | h_lm)
    cl = torch.zeros((B, Hl), **f32)
    cl2 = torch.zeros((B, Hl), **f32)
    hl = torch.zeros((B, Hl), **f32)
    hl_d = torch.empty((B, Hl), **f32) if lm_drop is not None else hl
    xh = torch.zeros((B, E + Hd), **f32)      # (xin | h_dec)
    xh2 = torch.zeros((B, E + Hd), **f32)
    c = torch.zeros((B, Hd), **f32)
    c2 = torch.zeros((B, Hd), **f32)
    cat = torch.zeros((B, Hd + D), **f32)     # (c_new | ctx)
    gl = torch.empty((B, 4 * Hl), **f32)
    gd = torch.empty((B, 4 * Hd), **f32)
    acts_l = torch.empty((B, 4 * Hl), **f32)
    acts_d = torch.empty((B, 4 * Hd), **f32)
    cn_l = torch.empty((B, Hl), **f32)
    y = torch.empty((B, A), **f32)
    alpha = torch.empty((B, Tn), **f32)
    nxt = torch.empty((B,), dtype=torch.int64, device=dev)
    big = torch.full((B,), 1 << 30, dtype=torch.int32, device=dev)   # LM-LSTM is never frozen
    for t in range(U):
        call("e2e_embed_gather", B, E, v["emb"], tok, u)
        xl[:, :E].copy_(u)
        ops.gemm(xl, v["lm_k"], bias=v["lm_b"], out=gl)
        # LM-LSTM step: new c -> cl2, new h -> xl[:, E:] (in place for the next step) and hl
        call("e2e_dec_pointwise_fwd", B, Hl, t, gl, cl, xl[:, E:], E + Hl, big, acts_l, cn_l, Hl, cl2,
             xl[:, E:], E + Hl, hl)
        cl, cl2 = cl2, cl
        if lm_drop is not None:     # the [U*B, Hl] mask of the training pass, rows t*B .. t*B+B
            call("e2e_dropout", B * Hl, hl, hl_d, float(lm_drop[0]), int(lm_drop[1]), int(lm_drop[2]), t * B * Hl,
                 ops.seed_dev(lm_drop[1]))
        m = ops.gemm(hl_d, v["sp_k"], bias=v["sp_b"]) if v["sp_k"] is not None else hl_d
        # xin = [m, ctx_prev] . in_k + in_b
        ops.gemm(m, v["in_k"][:Hd], bias=v["in_b"], out=xh[:, :E])
        ops.gemm(cat[:, Hd:], v["in_k"][Hd:], out=xh[:, :E], accumulate=True)
        ops.gemm(xh, v["dec_k"], bias=v["dec_b"], out=gd)
        call("e2e_dec_pointwise_fwd", B, Hd, t, gd, c, xh[:, E:], E + Hd, lens_i32, acts_d, cat, Hd + D, c2,
             xh2[:, E:], E + Hd, None)
        c, c2 = c2, c
        xh, xh2 = xh2, xh
        ops.gemm(cat[:, :Hd], v["q_k"], bias=v["q_b"], out=y)
        call("e2e_attn_fwd", B, Tn, Tp, A, D, HF, enc_flat, enc_len_i32, y, v["attn_v"], alpha, cat[:, Hd:], Hd + D)
        proj = ops.gemm(cat, v["ap_k"], bias=v["ap_b"])
        lg = logits[t * B:(t + 1) * B]
        ops.gemm(proj, v["out_k"], bias=v["out_b"], out=lg)
        if rule is None:
            call("e2e_argmax_rows", B, V, lg, V, nxt)
            tok = nxt
        else:
            tok = rule["ids"][t + 1]
            if rule["use_sample"][t + 1]:
                call("e2e_sample_rows", B, V, lg, V, int(rule["seed"]), int(rule["offset"]), (t + 1) * B, tok)
    call("e2e_mask_rows", U, B, V, logits, lens_i32)
    return logits, None

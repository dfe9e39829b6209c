"""Beam search over the attention decoder on the GPU, batched over utterances.

Same surface as the reference `BeamSearch` (beam_search.py:15-350):
`BeamSearch(ckpt_path, search_params)(encoder_hidden_states[T_enc, D]) -> ids`
(1-D int array, trailing EOS included, beam_search.py:338), same `class_params`.
`ckpt_path` may be a dict of weights keyed by the TF variable names of
beam_search.py:56-98, or the path of an .npz holding them (TF checkpoint I/O is
out of scope).

Where the reference loops utterance x step x hypothesis in NumPy, here every live
hypothesis of every utterance is one row of a float64 batch on the device
(e2e_*_f64 kernels keep the reference's dtype flow); only the
O(k^2) candidate merge per utterance -- `np.argpartition` over k*k scores,
back-pointers `idx // k`, EOS bookkeeping (beam_search.py:294-329) -- runs on the
host, on k-sized arrays, exactly as the reference does it.
"""
import numpy as np
import torch

from . import ops
from ._lib import call
from .base_params import BaseParams, Bunch
from .data_utils import EOS_ID, GO_ID
from .host_utils import BeamEntry  # noqa: F401  (record type kept for API parity)


class BeamSearch(BaseParams):
    """Implementation of beam search for the attention decoder."""

    MAX_STEPS = 120          # loop bound of beam_search.py:269

    @classmethod
    def class_params(cls):
        params = Bunch()
        params['beam_size'] = 4
        params['lm_weight'] = 0.0
        params['lm_path'] = ""
        params['word_ins_penalty'] = 0
        params['cov_penalty'] = 0.0          # parsed but unused in the reference too (beam_search.py:27,349)
        return params

    def __init__(self, ckpt_path, search_params=None, device="cuda", task="char"):
        self.device = torch.device(device)
        self.search_params = self.class_params() if search_params is None else search_params
        self.dec_params = self.map_dec_variables(self.get_model_params(ckpt_path), task)
        # the reference always loads "LM" weights too (beam_search.py:45-46); with lm_path pointing at the
        # same checkpoint they are the decoder's own LM-LSTM / projections
        lm_src = self.search_params.lm_path
        self.lm_params = self.dec_params if not lm_src or isinstance(lm_src, str) else \
            self.map_dec_variables(self.get_model_params(lm_src), task)
        self.use_lm = not (self.search_params.lm_path is None or self.search_params.lm_weight == 0.0)

    def get_model_params(self, ckpt_path):
        if isinstance(ckpt_path, dict):
            return ckpt_path
        return dict(np.load(ckpt_path))

    def map_dec_variables(self, var_dict, task="char"):
        """Name -> device tensor mapping (beam_search.py:53-109); float32 like the checkpoint."""
        pre = "model/rnn_decoder_%s/" % task
        names = dict(lm_lstm_w="rnn/basic_lstm_cell/kernel", lm_lstm_b="rnn/basic_lstm_cell/bias",
                     dec_lstm_w="rnn/basic_lstm_cell_1/kernel", dec_lstm_b="rnn/basic_lstm_cell_1/bias",
                     attn_dec_w="rnn/Attention/kernel", attn_dec_b="rnn/Attention/bias",
                     inp_w="rnn/InputProjection/kernel", inp_b="rnn/InputProjection/bias",
                     attn_proj_w="rnn/AttnProjection/kernel", attn_proj_b="rnn/AttnProjection/bias",
                     out_w="rnn/OutputProjection/kernel", out_b="rnn/OutputProjection/bias",
                     attn_v="AttnV", embedding="decoder/embedding")
        p = Bunch()
        for k, n in names.items():
            p[k] = torch.as_tensor(np.asarray(var_dict[pre + n], np.float32)).contiguous().to(self.device)
        p.attn_enc_w = torch.as_tensor(np.squeeze(np.asarray(var_dict[pre + "AttnW"], np.float32))).contiguous().to(self.device)
        if pre + "rnn/SimpleProjection/kernel" in var_dict:
            p.simple_w = torch.as_tensor(np.asarray(var_dict[pre + "rnn/SimpleProjection/kernel"], np.float32)).to(self.device)
            p.simple_b = torch.as_tensor(np.asarray(var_dict[pre + "rnn/SimpleProjection/bias"], np.float32)).to(self.device)
        else:
            p.simple_w = p.simple_b = None
        return p

    # ------------------------------------------------------------------
    def __call__(self, encoder_hidden_states):
        """Beam search for one utterance: [T_enc, D] (or [1, T_enc, D]) -> 1-D int array."""
        return self.decode_batch([encoder_hidden_states])[0]

    def _gemm64(self, a, w, b):
        out = torch.empty((a.shape[0], w.shape[1]), dtype=torch.float64, device=self.device)
        call("e2e_gemm_f64", a.shape[0], w.shape[1], a.shape[1], a, a.stride(0), w, w.stride(0), out, out.stride(0), b)
        return out

    def _lstm(self, x, c, h, w, b):
        """BasicLSTM step on rows: [x, h] . w + b -> (new_c, new_h)."""
        n, H = c.shape
        z = self._gemm64(torch.cat([x, h], dim=1), w, b)
        c2 = torch.empty_like(c)
        h2 = torch.empty_like(h)
        call("e2e_lstm_step_f64", n, H, z, c, c2, h2, H)
        return c2, h2

    def decode_batch(self, enc_list, return_scores=False):
        """Decode a list of utterances ([T_i, D] arrays) together; returns a list of id arrays."""
        sp, p, lp, dev = self.search_params, self.dec_params, self.lm_params, self.device
        beam = int(sp.beam_size)
        f64 = dict(dtype=torch.float64, device=dev)
        encs = []
        for e in enc_list:
            e = e.detach().cpu().numpy() if isinstance(e, torch.Tensor) else np.asarray(e)
            if e.ndim == 3:
                e = np.squeeze(e, axis=0)
            encs.append(np.ascontiguousarray(e, np.float32))
        N = len(encs)
        Ts = np.array([e.shape[0] for e in encs], np.int32)
        offs = np.concatenate([[0], np.cumsum(Ts)[:-1]]).astype(np.int32)
        enc_all = torch.from_numpy(np.concatenate(encs, axis=0)).to(dev)
        D = enc_all.shape[1]
        A = p.attn_enc_w.shape[1]
        V, E = p.embedding.shape
        Hd, Hl = p.dec_lstm_w.shape[1] // 4, p.lm_lstm_w.shape[1] // 4
        HF = ops.gemm(enc_all, p.attn_enc_w, mode=0)                      # float32 x float32 (beam_search.py:148)
        Tmax = int(Ts.max())

        # hypothesis rows (host bookkeeping): utterance, token sequence, model score
        utt = list(range(N))
        seqs = [[] for _ in range(N)]
        scores = [0.0] * N
        k_u = [beam] * N                  # current beam size per utterance (shrinks at EOS)
        final = [[] for _ in range(N)]
        tok = np.full(N, GO_ID, np.int64)
        st = dict(dc=torch.zeros((N, Hd), **f64), dh=torch.zeros((N, Hd), **f64),
                  lc=torch.zeros((N, Hl), **f64), lh=torch.zeros((N, Hl), **f64),
                  ctx=torch.zeros((N, D), **f64))
        if self.use_lm:
            st["mc"] = torch.zeros((N, Hl), **f64)
            st["mh"] = torch.zeros((N, Hl), **f64)
        step = 0
        while step < self.MAX_STEPS and len(utt) > 0:
            n = len(utt)
            tok_d = torch.from_numpy(tok).to(dev)
            x = torch.empty((n, E), **f64)
            call("e2e_embed_gather_f64", n, E, p.embedding, tok_d, x, E)
            # decoder's LM-LSTM, SimpleProjection, InputProjection, decoder LSTM (beam_search.py:182-191)
            lc, lh = self._lstm(x, st["lc"], st["lh"], p.lm_lstm_w, p.lm_lstm_b)
            m = lh if p.simple_w is None else self._gemm64(lh, p.simple_w, p.simple_b)
            x_dec = self._gemm64(torch.cat([m, st["ctx"]], dim=1), p.inp_w, p.inp_b)
            dc, dh = self._lstm(x_dec, st["dc"], st["dh"], p.dec_lstm_w, p.dec_lstm_b)
            # attention with the CELL state as query (beam_search.py:193), AttnProjection, OutputProjection
            y = self._gemm64(dc, p.attn_dec_w, p.attn_dec_b)
            ctx = torch.empty((n, D), **f64)
            uidx = np.asarray(utt)
            call("e2e_attn_beam_f64", n, A, D, Tmax, HF, enc_all, torch.from_numpy(offs[uidx]).to(dev),
                 torch.from_numpy(Ts[uidx]).to(dev), y, p.attn_v, ctx, D)
            proj = self._gemm64(torch.cat([dc, ctx], dim=1), p.attn_proj_w, p.attn_proj_b)
            logits = self._gemm64(proj, p.out_w, p.out_b)
            lm_logits = None
            if self.use_lm:                                                # LM branch (beam_search.py:200-207)
                x_lm = torch.empty((n, E), **f64)
                call("e2e_embed_gather_f64", n, E, lp.embedding, tok_d, x_lm, E)
                mc, mh = self._lstm(x_lm, st["mc"], st["mh"], lp.lm_lstm_w, lp.lm_lstm_b)
                lo = mh if lp.simple_w is None else self._gemm64(mh, lp.simple_w, lp.simple_b)
                lm_logits = self._gemm64(lo, lp.out_w, lp.out_b)
            krow = np.array([k_u[u] for u in utt], np.int32)
            out_idx = torch.empty((n, beam), dtype=torch.int32, device=dev)
            out_val = torch.empty((n, beam), **f64)
            scratch = torch.empty((n, V), **f64)
            call("e2e_logsoftmax_topk_f64", n, V, logits, lm_logits, float(sp.lm_weight),
                 torch.from_numpy(krow).to(dev), beam, out_idx, out_val, scratch)
            idx_h = out_idx.cpu().numpy()
            val_h = out_val.cpu().numpy()
            # ---- host: merge candidates per utterance (beam_search.py:294-329)
            new_utt, new_seqs, new_scores, new_tok, parents = [], [], [], [], []
            r = 0
            while r < n:
                u = utt[r]
                r1 = r
                while r1 < n and utt[r1] == u:
                    r1 += 1
                k = k_u[u]
                if step == 0:                                              # single GO hypothesis (:255-266)
                    cand_scores = val_h[r, :k]
                    cand_tokens = idx_h[r, :k]
                    sel = np.arange(k)
                    par = np.zeros(k, np.int64)
                    model_scores = cand_scores
                else:
                    all_scores = np.concatenate([val_h[i, :k] + scores[i] for i in range(r, r1)])
                    cand_tokens = np.concatenate([idx_h[i, :k] for i in range(r, r1)])
                    sel = np.argpartition(all_scores, -k)[-k:]
                    par = sel // k
                    model_scores = all_scores
                for j in range(k):                                         # bound fixed before k shrinks (:310)
                    pr = r + int(par[j])
                    t_new = int(cand_tokens[sel[j]])
                    seq = seqs[pr] + [t_new]
                    sc = float(model_scores[sel[j]]) + sp.word_ins_penalty * len(seq)
                    if t_new == EOS_ID:
                        final[u].append((seq, sc))
                        k_u[u] -= 1
                    else:
                        new_utt.append(u); new_seqs.append(seq); new_scores.append(sc)
                        new_tok.append(t_new); parents.append(pr)
                r = r1
            step += 1
            if step >= self.MAX_STEPS or not new_utt:
                # leftovers join the final list (beam_search.py:332)
                for u, seq, sc in zip(new_utt, new_seqs, new_scores):
                    final[u].append((seq, sc))
                break
            sel_rows = torch.from_numpy(np.asarray(parents, np.int64)).to(dev)
            st = dict(dc=dc.index_select(0, sel_rows), dh=dh.index_select(0, sel_rows),
                      lc=lc.index_select(0, sel_rows), lh=lh.index_select(0, sel_rows),
                      ctx=ctx.index_select(0, sel_rows))
            if self.use_lm:
                st["mc"] = mc.index_select(0, sel_rows)
                st["mh"] = mh.index_select(0, sel_rows)
            utt, seqs, scores, tok = new_utt, new_seqs, new_scores, np.asarray(new_tok, np.int64)
        outs, outs_sc = [], []
        for u in range(N):
            best = max(final[u], key=lambda e: e[1])                        # first maximum, no length norm (:336)
            outs.append(np.stack(best[0], axis=0))
            outs_sc.append(best[1])
        return (outs, outs_sc) if return_scores else outs

    @classmethod
    def add_parse_options(cls, parser):
        # flag names and defaults of beam_search.py:340-350
        parser.add_argument("-beam_size", default=1, type=int, help="Beam size")
        parser.add_argument("-lm_weight", default=0.0, type=float, help="LM weight in decoding")
        parser.add_argument("-lm_path", default="", type=str, help="LM ckpt path")
        parser.add_argument("-cov_penalty", default=0.0, type=float, help="Coverage penalty")

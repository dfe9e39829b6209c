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
host as the reference does it, vectorised over utterances (`merge_candidates`).
"""
import numpy as np
import torch

from . import ops
from ._lib import call
from .base_params import BaseParams, Bunch
from .data_utils import EOS_ID, GO_ID
from .host_utils import BeamEntry  # noqa: F401  (record type kept for API parity)


def merge_candidates(utt, scores, k_u, idx_h, val_h, step, word_ins_penalty):
    """Candidate merge of one step for every utterance at once (beam_search.py:255-266 for the GO step, :294-329
    after it), vectorised over utterances with NumPy.

    utt [n] (ascending): utterance of each hypothesis row; scores [n]: its score; k_u [N]: current beam size per
    utterance; idx_h / val_h [n, beam]: per-row top tokens and their log-probabilities (entries >= k_u are unused).
    Per utterance the reference concatenates `val[i, :k] + score[i]` over its rows, takes `np.argpartition(., -k)[-k:]`
    IN THAT ORDER, derives back-pointers `idx // k`, adds `word_ins_penalty * len(seq)` and retires EOS candidates
    (the beam shrinks).  Utterances with the same (rows, k) are stacked and partitioned row-wise -- the same
    introselect on the same 1-D data, so the selection and its order are the reference's.

    Returns (rows, finished): rows = dict(utt, parent, tok, score) of the surviving hypotheses ordered by
    (utterance, candidate position); finished = list of (utterance, parent row, score) of candidates that emitted
    EOS, in the reference's append order within each utterance.  k_u is decremented in place."""
    n = len(utt)
    starts = np.flatnonzero(np.r_[True, utt[1:] != utt[:-1]]) if n else np.zeros(0, np.int64)
    counts = np.diff(np.r_[starts, n])
    us = utt[starts]
    ks = k_u[us]
    parts = []
    for c, k in sorted(set(zip(counts.tolist(), ks.tolist()))):
        if k <= 0:
            continue
        g = np.flatnonzero((counts == c) & (ks == k))
        rows = starts[g][:, None] + np.arange(c)[None, :]                    # [G, c]
        flat = (val_h[rows][:, :, :k] + scores[rows][:, :, None]).reshape(len(g), c * k)
        toks = idx_h[rows][:, :, :k].reshape(len(g), c * k)
        if step == 0:                                                        # single GO row: candidates in order
            sel = np.tile(np.arange(k), (len(g), 1))
        else:
            sel = np.argpartition(flat, -k, axis=1)[:, -k:]
        par = starts[g][:, None] + sel // k
        parts.append((np.repeat(us[g], k), np.tile(np.arange(k), len(g)), par.reshape(-1),
                      np.take_along_axis(toks, sel, 1).reshape(-1).astype(np.int64),
                      # the step-0 candidates carry the bare model score (beam_search.py:258-260); later ones add
                      # word_ins_penalty * len(new_index_seq) (:321-322)
                      np.take_along_axis(flat, sel, 1).reshape(-1) + (word_ins_penalty * (step + 1) if step else 0)))
    if not parts:
        e = np.zeros(0, np.int64)
        return dict(utt=e, parent=e, tok=e, score=np.zeros(0)), []
    u_all, j_all, par_all, tok_all, sc_all = [np.concatenate(x) for x in zip(*parts)]
    order = np.lexsort((j_all, u_all))
    u_all, par_all, tok_all, sc_all = u_all[order], par_all[order], tok_all[order], sc_all[order]
    eos = tok_all == EOS_ID
    finished = [(int(u), int(p_), float(s_)) for u, p_, s_ in zip(u_all[eos], par_all[eos], sc_all[eos])]
    np.subtract.at(k_u, u_all[eos], 1)
    keep = ~eos
    return dict(utt=u_all[keep], parent=par_all[keep], tok=tok_all[keep], score=sc_all[keep]), finished


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
        self.use_lm = not (self.search_params.lm_path is None or self.search_params.lm_weight == 0.0)
        self.lm_params = self._load_lm(self.search_params.lm_path, ckpt_path, task)

    def _load_lm(self, lm_src, ckpt_path, task):
        """`map_lm_variables(get_model_params(lm_path))` (beam_search.py:45-46,111-134): the LM-LSTM, SimpleProjection,
        OutputProjection and embedding of the checkpoint at `lm_path` (a dict of weights is accepted like for
        `ckpt_path`).  The same checkpoint (or none, while lm_weight == 0 leaves the LM branch unused) reuses the
        decoder's tensors; a path that cannot be read is an error whenever the LM is used."""
        import os
        if isinstance(lm_src, dict):
            return self.dec_params if lm_src is ckpt_path else self.map_dec_variables(lm_src, task, lm_only=True)
        if not lm_src or (isinstance(ckpt_path, str) and lm_src == ckpt_path):
            return self.dec_params
        from .tf_checkpoint import checkpoint_exists
        if checkpoint_exists(lm_src) or os.path.exists(lm_src):
            return self.map_dec_variables(self.get_model_params(lm_src), task, lm_only=True)
        if self.use_lm:
            raise FileNotFoundError("BeamSearch: lm_path %r is not a readable checkpoint (lm_weight = %g)"
                                    % (lm_src, self.search_params.lm_weight))
        return self.dec_params

    def get_model_params(self, ckpt_path):
        """Weights by TF variable name (beam_search.py:36-47 reads them with tf.train.NewCheckpointReader): a dict, a
        TF V2 checkpoint prefix (`<prefix>.index` + data shards, read by tf_checkpoint.py) or an .npz file."""
        if isinstance(ckpt_path, dict):
            return ckpt_path
        from .tf_checkpoint import checkpoint_exists, read_checkpoint
        if checkpoint_exists(ckpt_path):
            return read_checkpoint(ckpt_path)
        return dict(np.load(ckpt_path))

    def map_dec_variables(self, var_dict, task="char", lm_only=False):
        """Name -> device tensor mapping (beam_search.py:53-109); float32 like the checkpoint.  lm_only: the subset
        map_lm_variables reads (beam_search.py:111-134), under the decoder's field names."""
        pre = "model/rnn_decoder_%s/" % task
        names = dict(lm_lstm_w="rnn/basic_lstm_cell/kernel", lm_lstm_b="rnn/basic_lstm_cell/bias",
                     out_w="rnn/OutputProjection/kernel", out_b="rnn/OutputProjection/bias",
                     embedding="decoder/embedding") if lm_only else dict(lm_lstm_w="rnn/basic_lstm_cell/kernel", lm_lstm_b="rnn/basic_lstm_cell/bias",
                     dec_lstm_w="rnn/basic_lstm_cell_1/kernel", dec_lstm_b="rnn/basic_lstm_cell_1/bias",
                     attn_dec_w="rnn/Attention/kernel", attn_dec_b="rnn/Attention/bias",
                     inp_w="rnn/InputProjection/kernel", inp_b="rnn/InputProjection/bias",
                     attn_proj_w="rnn/AttnProjection/kernel", attn_proj_b="rnn/AttnProjection/bias",
                     out_w="rnn/OutputProjection/kernel", out_b="rnn/OutputProjection/bias",
                     attn_v="AttnV", embedding="decoder/embedding")
        p = Bunch()
        for k, n in names.items():
            p[k] = torch.as_tensor(np.asarray(var_dict[pre + n], np.float32)).contiguous().to(self.device)
        if not lm_only:
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

        # hypothesis rows (host bookkeeping, NumPy arrays): utterance, token history, model score
        utt = np.arange(N)
        hist = np.zeros((N, 0), np.int64)     # tokens emitted so far, one row per hypothesis
        scores = np.zeros(N, np.float64)
        k_u = np.full(N, beam, np.int64)      # current beam size per utterance (shrinks at EOS)
        final = [[] for _ in range(N)]
        tok = np.full(N, GO_ID, np.int64)
        st = dict(dc=torch.zeros((N, Hd), **f64), dh=torch.zeros((N, Hd), **f64),
                  lc=torch.zeros((N, Hl), **f64), lh=torch.zeros((N, Hl), **f64),
                  ctx=torch.zeros((N, D), **f64))
        if self.use_lm:
            Hm = lp.lm_lstm_w.shape[1] // 4
            st["mc"] = torch.zeros((N, Hm), **f64)
            st["mh"] = torch.zeros((N, Hm), **f64)
        step = 0
        while step < self.MAX_STEPS and len(utt) > 0:
            n = len(utt)
            tok_d = torch.from_numpy(np.ascontiguousarray(tok)).to(dev)
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
            uidx = utt
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
            krow = k_u[utt].astype(np.int32)
            out_idx = torch.empty((n, beam), dtype=torch.int32, device=dev)
            out_val = torch.empty((n, beam), **f64)
            scratch = torch.empty((n, V), **f64)
            call("e2e_logsoftmax_topk_f64", n, V, logits, lm_logits, float(sp.lm_weight),
                 torch.from_numpy(krow).to(dev), beam, out_idx, out_val, scratch)
            idx_h = out_idx.cpu().numpy()
            val_h = out_val.cpu().numpy()
            # ---- host: merge candidates per utterance (beam_search.py:294-329), all utterances at once
            new, finished = merge_candidates(utt, scores, k_u, idx_h, val_h, step, sp.word_ins_penalty)
            for u, pr, sc in finished:
                final[u].append((np.append(hist[pr], EOS_ID), sc))
            step += 1
            parents = new["parent"]
            hist = np.concatenate([hist[parents], new["tok"][:, None]], axis=1)
            if step >= self.MAX_STEPS or len(parents) == 0:
                # leftovers join the final list (beam_search.py:332)
                for u, seq, sc in zip(new["utt"], hist, new["score"]):
                    final[int(u)].append((seq, float(sc)))
                break
            sel_rows = torch.from_numpy(parents).to(dev)
            st = dict(dc=dc.index_select(0, sel_rows), dh=dh.index_select(0, sel_rows),
                      lc=lc.index_select(0, sel_rows), lh=lh.index_select(0, sel_rows),
                      ctx=ctx.index_select(0, sel_rows))
            if self.use_lm:
                st["mc"] = mc.index_select(0, sel_rows)
                st["mh"] = mh.index_select(0, sel_rows)
            utt, scores, tok = new["utt"], new["score"], new["tok"]
        outs, outs_sc = [], []
        for u in range(N):
            best = max(final[u], key=lambda e: e[1])                        # first maximum, no length norm (:336)
            outs.append(np.asarray(best[0], np.int64))
            outs_sc.append(best[1])
        return (outs, outs_sc) if return_scores else outs

    @classmethod
    def add_parse_options(cls, parser):
        # flag names and defaults of beam_search.py:340-350
        parser.add_argument("-beam_size", default=1, type=int, help="Beam size")
        parser.add_argument("-lm_weight", default=0.0, type=float, help="LM weight in decoding")
        parser.add_argument("-lm_path", default="", type=str, help="LM ckpt path")
        parser.add_argument("-cov_penalty", default=0.0, type=float, help="Coverage penalty")

"""Host side of the batched beam search without a GPU: `decode_batch` runs with every C-ABI kernel replaced by a NumPy
emulation that follows the kernel's contract in include/e2e_asr_b200.h (same arguments, same in-place outputs), so the
plumbing around the kernels -- slot-state buffers, the token table and gate-interleaved LSTM weights, the operands taken
in place instead of concatenated, the merge writing into the slot state, the back-pointer walk -- is checked against the
oracle's ids on the CPU.  The kernels themselves are checked on the GPU (tests/test_gpu_beam.py)."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from e2e_asr_b200 import beam_search as bsm
from e2e_asr_b200 import synth
from e2e_asr_b200.data_utils import EOS_ID
from oracle import beam as ob

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import gen_golden as gg  # noqa: E402


def _np(t):
    return None if t is None else t.numpy()


def _view(ptr, n, ctype, dtype):
    return np.frombuffer((ctype * n).from_address(ptr), dtype=dtype)


def _strided(t, rows, cols, ld):
    """[rows, cols] view with row stride ld of a tensor passed by its first element."""
    a = t.numpy().reshape(-1) if t.is_contiguous() else None
    if a is None:                              # a column / row slice: rebuild from storage
        base = np.frombuffer((ctypes.c_double * (ld * (rows - 1) + cols)).from_address(t.data_ptr()), np.float64)
        return np.lib.stride_tricks.as_strided(base, (rows, cols), (ld * 8, 8))
    return np.lib.stride_tricks.as_strided(a, (rows, cols), (ld * a.itemsize, a.itemsize))


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


def _lstm(g, c_prev, H):
    cn = c_prev * _sig(g[:, 2 * H:3 * H] + 1.0) + _sig(g[:, :H]) * np.tanh(g[:, H:2 * H])
    return cn, _sig(g[:, 3 * H:]) * np.tanh(cn)


class Emu:
    """NumPy stand-ins for the kernels decode_batch launches."""

    def __init__(self):
        self.names = []

    def __call__(self, name, *a, **kw):
        self.names.append(name)
        return getattr(self, name)(*a)

    def _cat(self, M, K1, K2, A1, lda1, A2, lda2, B, ldb, bias, Z, ldz, zrow):
        a = _strided(A1, M, K1, lda1)
        if K2:
            a = np.concatenate([a, _strided(A2, M, K2, lda2)], axis=1)
        N = B.shape[1]
        out = a @ _strided(B, K1 + K2, N, ldb)
        if bias is not None:
            out = out + bias.numpy().astype(np.float64)
        if Z is not None:
            out = out + _strided(Z, Z.shape[0], N, ldz)[zrow.numpy()[:M]]
        return out

    def e2e_gemm_f64d(self, M, N, K, A, lda, B, ldb, C, ldc, bias):
        _strided(C, M, N, ldc)[:] = self._cat(M, K, 0, A, lda, None, 0, B, ldb, bias, None, 0, None)

    def e2e_gemm_f64d_cat(self, M, N, K1, K2, A1, lda1, A2, lda2, B, ldb, C, ldc, bias, Z, ldz, zrow):
        assert B.shape[1] == N and K1 % 16 == 0 and K2 % 16 == 0
        _strided(C, M, N, ldc)[:] = self._cat(M, K1, K2, A1, lda1, A2, lda2, B, ldb, bias, Z, ldz, zrow)

    def e2e_gemm_f64d_lstm(self, M, H, K1, K2, A1, lda1, A2, lda2, B, ldb, bias, Z, ldz, zrow, c_prev, c_out, h_out,
                           ldh):
        g = self._cat(M, K1, K2, A1, lda1, A2, lda2, B, ldb, bias, Z, ldz, zrow)
        g = g[:, np.argsort(bsm.lstm_gate_perm(H))]              # interleaved columns -> i | j | f | o
        cn, hn = _lstm(g, c_prev.numpy(), H)
        c_out.numpy()[:] = cn
        _strided(h_out, M, H, ldh)[:] = hn

    def e2e_gemm_f64(self, M, N, K, A, lda, B, ldb, C, ldc, bias):
        _strided(C, M, N, ldc)[:] = (_strided(A, M, K, lda) @ B.numpy().astype(np.float64)
                                     + (0.0 if bias is None else bias.numpy().astype(np.float64)))

    def e2e_lstm_step_f64(self, n, H, z, c_prev, c_out, h_out, ldh):
        cn, hn = _lstm(z.numpy(), c_prev.numpy(), H)
        c_out.numpy()[:] = cn
        _strided(h_out, n, H, ldh)[:] = hn

    def e2e_embed_gather_f64(self, n, E, emb, ids, out, ldo):
        _strided(out, n, E, ldo)[:] = emb.numpy()[ids.numpy()[:n]].astype(np.float64)

    def e2e_exp2x_f64(self, n, x, out):
        out.numpy().reshape(-1)[:n] = np.exp(np.clip(2.0 * x.numpy().reshape(-1)[:n].astype(np.float64), -300, 300))

    def _attn(self, N, beam, A, D, HF64, enc, row_off, Tlen, y, v, ctx, from_exp):
        ro, tl = row_off.numpy(), Tlen.numpy()
        for r in range(N * beam):
            sl = slice(int(ro[r]), int(ro[r]) + int(tl[r]))
            if from_exp:
                t = 1.0 - 2.0 / (HF64[sl] * np.exp(np.clip(2.0 * y.numpy()[r], -300, 300)) + 1.0)
            else:
                t = np.tanh(HF64[sl] + y.numpy()[r])
            s = t @ v.numpy().astype(np.float64)
            e = np.exp(s - s.max())
            ctx.numpy()[r, :D] = (e / e.sum()) @ enc.numpy()[sl].astype(np.float64)

    def e2e_attn_beam_group_e_f64(self, N, beam, A, D, Tmax, EHF, enc, row_off, Tlen, y, v, ctx, ldctx):
        self._attn(N, beam, A, D, EHF.numpy(), enc, row_off, Tlen, y, v, ctx, True)

    def e2e_attn_beam_group_f64(self, N, beam, A, D, Tmax, HF, enc, row_off, Tlen, y, v, ctx, ldctx):
        self._attn(N, beam, A, D, HF.numpy().astype(np.float64), enc, row_off, Tlen, y, v, ctx, False)

    def e2e_logsoftmax_topk_f64(self, n, V, logits, lm_logits, lm_weight, krow, kmax, out_idx, out_val, scratch):
        def logsm(x):
            e = np.exp(x - x.max(axis=1, keepdims=True))
            return np.log(e / e.sum(axis=1, keepdims=True))
        comb = logsm(logits.numpy())
        if lm_logits is not None:
            comb = comb + lm_weight * logsm(lm_logits.numpy())
        oi, ov, kr = out_idx.numpy(), out_val.numpy(), krow.numpy()
        for r in range(n):
            order = np.lexsort((np.arange(V), -comb[r]))        # descending score, ties to the lower index
            oi[r], ov[r] = -1, -np.inf
            k = int(kr[r])
            oi[r, :k], ov[r, :k] = order[:k], comb[r, order[:k]]

    def e2e_beam_merge(self, a):
        N, beam, R = a.N, a.beam, a.R
        i32 = lambda p, n: _view(p, n, ctypes.c_int32, np.int32)
        f64 = lambda p, n: _view(p, n, ctypes.c_double, np.float64)
        step = int(i32(a.step, 1)[0])
        out_idx, out_val = i32(a.out_idx, R * beam).reshape(R, beam), f64(a.out_val, R * beam).reshape(R, beam)
        score, alive, k_u = f64(a.score, R), i32(a.alive, R), i32(a.k_u, N)
        new_tok = _view(a.new_tok, R, ctypes.c_int64, np.int64)
        new_score, parent, new_alive, krow = f64(a.new_score, R), i32(a.parent, R), i32(a.new_alive, R), i32(a.krow, R)
        par_hist, tok_hist = i32(a.par_hist + 4 * step * R, R), i32(a.tok_hist + 4 * step * R, R)
        fin_cnt, fin_step, fin_row, fin_score = i32(a.fin_cnt, N), i32(a.fin_step, R), i32(a.fin_row, R), f64(a.fin_score, R)
        n_live = i32(a.n_live, 1)
        for u in range(N):
            k, base = int(k_u[u]), u * beam
            cand = []                                                # gathered BEFORE any write (the outputs may alias)
            for slot in range(beam):
                row = base + slot
                if alive[row]:
                    cand += [(out_val[row, j] + score[row], int(out_idx[row, j]), row) for j in range(k)]
            pen = a.word_ins_penalty * (step + 1) if step > 0 else 0.0
            taken, n_new, k_left, nfin = set(), 0, k, int(fin_cnt[u])
            for _ in range(k):
                best = None
                for c, (v, tok, prow) in enumerate(cand):
                    if c not in taken and tok >= 0 and (best is None or v > cand[best][0]):
                        best = c
                if best is None:
                    break
                taken.add(best)
                v, tok, prow = cand[best]
                if tok == a.eos_id:
                    if nfin < beam:
                        fin_step[base + nfin], fin_row[base + nfin], fin_score[base + nfin] = step, prow, v + pen
                    nfin += 1
                    k_left -= 1
                else:
                    row = base + n_new
                    new_tok[row], new_score[row], parent[row] = tok, v + pen, prow
                    par_hist[row], tok_hist[row] = prow, tok
                    n_new += 1
            for slot in range(n_new, beam):
                row = base + slot
                new_tok[row], new_score[row], parent[row], par_hist[row], tok_hist[row] = 0, 0.0, base, -1, -1
            for slot in range(beam):
                new_alive[base + slot] = 1 if slot < n_new else 0
                krow[base + slot] = k_left if slot < n_new else 0
            k_u[u], fin_cnt[u] = k_left, nfin
            if k_left > 0:
                n_live[0] += k_left

    def e2e_beam_gather(self, R, parent, g):
        par = parent.numpy()
        for m in range(g.nmat):
            w = g.width[m]
            src = _view(g.src[m], R * w, ctypes.c_double, np.float64).reshape(R, w)
            dst = _view(g.dst[m], R * w, ctypes.c_double, np.float64).reshape(R, w)
            dst[:] = src[par].copy()


@pytest.fixture
def emulated(monkeypatch):
    emu = Emu()

    def call(name, *args, **kw):
        conv = [ctypes.cast(ctypes.pointer(x), ctypes.POINTER(type(x))).contents if isinstance(x, ctypes.Structure) else x
                for x in args]
        return emu(name, *conv)

    monkeypatch.setattr(bsm, "call", call)
    monkeypatch.setattr(bsm.ops, "gemm", lambda a, b, mode=0, out=None: out.copy_(a @ b))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    return emu


FLAGS = [dict(), dict(fast_step=False), dict(token_table=False), dict(fused_lstm=False), dict(exp_attention=False),
         dict(token_table=False, fused_lstm=False, exp_attention=False)]


@pytest.mark.parametrize("hl32", [False, True], ids=["Hl16", "Hl32"])
@pytest.mark.parametrize("flags", FLAGS, ids=["default"] + ["-".join(sorted(f)) for f in FLAGS[1:]])
def test_decode_batch_host_logic_with_emulated_kernels(emulated, flags, hl32):
    # widths in multiples of 16 (the tensor-core step's tiles); Hl != Hd adds the SimpleProjection
    cfg = synth.get_config("tiny", H=8, E=16, A=16, Hd=16, Hl=32 if hl32 else 16)
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    rng = np.random.Generator(np.random.PCG64(3))
    encs = [(np.tanh(rng.standard_normal((int(rng.integers(5, 12)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(3)]
    for k, lmw in ((3, 0.0), (2, 0.3)):
        sp = bsm.BeamSearch.class_params()
        sp.beam_size, sp.lm_weight, sp.lm_path = k, lmw, w
        bs = bsm.BeamSearch(w, sp, device="cpu")
        for name, val in flags.items():
            setattr(bs, name, val)
        out, sc = bs.decode_batch(encs, return_scores=True, use_graph=False)
        for u, enc in enumerate(encs):
            ref, rs = ob.beam_search(w, enc, beam_size=k, lm_weight=lmw, return_score=True)
            np.testing.assert_array_equal(out[u], ref)
            assert out[u][-1] == EOS_ID or len(out[u]) == bs.MAX_STEPS
            assert abs(sc[u] - rs) <= 1e-6 * max(1.0, abs(rs))
    launched = set(emulated.names)
    assert next(iter(bs._plans.values())).fast == flags.get("fast_step", True)
    if flags.get("fast_step", True):
        assert "e2e_gemm_f64d_cat" in launched
        assert ("e2e_gemm_f64d_lstm" in launched) == flags.get("fused_lstm", True)
        assert ("e2e_attn_beam_group_e_f64" in launched) == flags.get("exp_attention", True)
        assert ("e2e_embed_gather_f64" in launched) == (not flags.get("token_table", True))


def test_widths_the_tensor_core_step_cannot_tile_fall_back(emulated):
    cfg = synth.get_config("tiny_b")                       # E = 12, Hl = 8: not multiples of 16
    w = gg.dec_weights(cfg, 12, 8.0, 2.0)
    rng = np.random.Generator(np.random.PCG64(4))
    encs = [(np.tanh(rng.standard_normal((7 + u, 2 * cfg.H))) * 0.8).astype(np.float32) for u in range(2)]
    sp = bsm.BeamSearch.class_params()
    sp.beam_size = 3
    bs = bsm.BeamSearch(w, sp, device="cpu")
    out = bs.decode_batch(encs, use_graph=False)
    assert not next(iter(bs._plans.values())).fast and "e2e_gemm_f64d_cat" not in emulated.names
    for u, enc in enumerate(encs):
        np.testing.assert_array_equal(out[u], ob.beam_search(w, enc, beam_size=3))


def test_device_resident_encoder_states_are_gathered_in_place(emulated):
    """Tensors that already live on the search's device skip the host staging; [1, T, D] inputs are accepted."""
    cfg = synth.get_config("tiny", H=8, E=16, A=16, Hd=16, Hl=16)
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    rng = np.random.Generator(np.random.PCG64(9))
    encs = [(np.tanh(rng.standard_normal((6 + 2 * u, 2 * cfg.H))) * 0.8).astype(np.float32) for u in range(3)]
    sp = bsm.BeamSearch.class_params()
    sp.beam_size = 3
    bs = bsm.BeamSearch(w, sp, device="cpu")
    as_t = [torch.from_numpy(encs[0]), torch.from_numpy(encs[1])[None], torch.from_numpy(encs[2]).double()]
    out_t = bs.decode_batch(as_t, use_graph=False)
    out_n = bs.decode_batch(encs, use_graph=False)
    for a, b, enc in zip(out_t, out_n, encs):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, ob.beam_search(w, enc, beam_size=3))

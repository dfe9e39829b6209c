"""Kernel-level parity on the GPU (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest
import torch

from e2e_asr_b200 import ops, synth
from oracle import model as om

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), dtype=dtype, device=DEV)


def relerr(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (3, 5, 7), (64, 1024, 512), (130, 257, 33), (7680, 1000, 256),
                                   (64, 256, 512), (300, 200, 4100)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm(M, N, K, ta, tb):
    rng = np.random.default_rng(M * 1000 + N + K)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    z = rng.standard_normal((M, N)).astype(np.float32)
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    out = ops.gemm(T(a), T(b), ta=ta, tb=tb)
    assert relerr(out, ref) < 2e-5
    out = ops.gemm(T(a), T(b), ta=ta, tb=tb, bias=T(bias), z=T(z))
    assert relerr(out, ref + bias + z) < 2e-5
    c = T(c0)
    ops.gemm(T(a), T(b), ta=ta, tb=tb, out=c, accumulate=True)
    assert relerr(c, ref + c0) < 2e-5


def test_gemm_strided_views():
    rng = np.random.default_rng(5)
    big = T(rng.standard_normal((50, 40)))
    w = T(rng.standard_normal((16, 24)))
    out = torch.zeros((50, 30), device=DEV)
    ops.gemm(big[:, 8:24], w, out=out[:, 3:27])
    ref = big[:, 8:24].cpu().numpy() @ w.cpu().numpy()
    assert relerr(out[:, 3:27], ref) < 2e-5
    assert float(out[:, :3].abs().max()) == 0 and float(out[:, 27:].abs().max()) == 0
    cs = ops.colsum(big[:, 8:24])
    assert relerr(cs, big[:, 8:24].cpu().numpy().sum(0)) < 2e-5


@pytest.mark.parametrize("rec_mode", [0, 1, 5, 6, 7, 8])
@pytest.mark.parametrize("B,T_,I,H", [(3, 11, 6, 8), (5, 20, 12, 16), (17, 9, 40, 24), (4, 30, 40, 256),
                                      (70, 6, 16, 32), (33, 25, 24, 128), (64, 40, 16, 256), (130, 7, 8, 256),
                                      (20, 9, 12, 512), (130, 5, 8, 512), (16, 33, 8, 512)])
def test_bilstm_layer_fwd_bwd(B, T_, I, H, rec_mode):
    """rec_mode 0 = fastest eligible recurrence, 1 = L2-exchange kernel, 5 / 6 = the warp-specialised cluster kernel with
    1 / 2 interleaved batch slices per cluster, 7 / 8 = with the forward / backward pass on the tf32 + bf16 products
    instead of the fp16 split scheme (backward: per-row scaled dz tiles).
    H = 512 under mode 0 runs the kernels of lstm_rec_h512.cu (W_hh hi plane in registers, lo plane in shared memory;
    130 rows = 18 clusters of 16, more than are co-resident)."""
    if rec_mode >= 5 and H not in (128, 256):
        pytest.skip("the warp-specialised cluster recurrence serves H in {128, 256}")
    from e2e_asr_b200 import _lib
    _lib.lib().e2e_set_rec_mode(rec_mode)
    try:
        _bilstm_case(B, T_, I, H)
    finally:
        _lib.lib().e2e_set_rec_mode(0)


def _bilstm_case(B, T_, I, H):
    rng = np.random.default_rng(B + T_ + I + H)
    lens = rng.integers(1, T_ + 1, size=B)
    lens[0] = T_
    x = rng.standard_normal((B, T_, I)).astype(np.float32)
    for b in range(B):
        x[b, lens[b]:] = 0
    # weight scale keeps the recurrence non-chaotic (gain ~1.2, the reference's U(+-0.075) at
    # H=256): in a chaotic regime any rounding difference is amplified exponentially with t
    ws = min(0.3, 1.2 / np.sqrt(H))
    ks = [rng.uniform(-ws, ws, (I + H, 4 * H)).astype(np.float32) for _ in range(2)]
    bs = [rng.uniform(-0.3, 0.3, (4 * H,)).astype(np.float32) for _ in range(2)]
    ref_out, cache = om.birnn_layer_fwd(x.astype(np.float64), lens, ks[0].astype(np.float64), bs[0].astype(np.float64),
                                        ks[1].astype(np.float64), bs[1].astype(np.float64))
    dout = rng.standard_normal(ref_out.shape)
    ref_dx, ref_g = om.birnn_layer_bwd(dout, cache)
    Tp = T_ + 2 - (T_ % 2)
    xp = torch.zeros((B, Tp, I), device=DEV)
    xp[:, :T_] = T(x)
    xp.requires_grad_(True)
    params = [T(ks[0]).requires_grad_(), T(bs[0]).requires_grad_(), T(ks[1]).requires_grad_(), T(bs[1]).requires_grad_()]
    out = ops.BiLSTMLayerFn.apply(xp, *params, T(lens, torch.int32), int(lens.max()))
    assert relerr(out[:, :T_], ref_out) < 1e-5
    assert float(out[:, T_:].abs().max()) == 0.0
    dp = torch.zeros_like(out)
    dp[:, :T_] = T(dout)
    out.backward(dp)
    ops.check_device_errors(DEV)
    assert relerr(xp.grad[:, :T_], ref_dx) < 2e-5
    for d in range(2):
        assert relerr(params[2 * d].grad, ref_g[d][0]) < 2e-5
        assert relerr(params[2 * d + 1].grad, ref_g[d][1]) < 2e-5


@pytest.mark.parametrize("impl", ["persist", "persist-gridbarrier", "loop"])
@pytest.mark.parametrize("cname", ["tiny", "tiny_b", "cfg1"])
def test_attn_decoder_fwd_bwd(cname, impl):
    """persist: the persistent decoder kernels with point-to-point row-block counters (default) / with three grid
    barriers per step; loop: the per-step kernels."""
    from e2e_asr_b200 import _lib
    ops.set_decoder_impl("loop" if impl == "loop" else "persist")
    _lib.lib().e2e_set_dec_sync(0 if impl == "persist-gridbarrier" else 1)
    try:
        _attn_decoder_case(cname)
    finally:
        ops.set_decoder_impl("persist")
        _lib.lib().e2e_set_dec_sync(1)


def _attn_decoder_case(cname):
    from e2e_asr_b200.attn_decoder import AttnDecoder
    from e2e_asr_b200.testing import model_params
    from e2e_asr_b200.variables import VariableStore
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    rng = np.random.default_rng(3)
    B, Tn, D = cfg.B, 7, 2 * cfg.H
    enc = np.tanh(rng.standard_normal((B, Tn, D)))
    enc_len = rng.integers(2, Tn + 1, size=B)
    enc_len[0] = Tn
    for b in range(B):
        enc[b, enc_len[b]:] = 0
    dec_inp = batch["char"].T
    seq_len = batch["char_len"]
    W64 = {k: v.astype(np.float64) for k, v in w.items()}
    ref_logits, cache = om.attn_decoder_fwd(W64, "char", dec_inp, seq_len, enc, enc_len)
    targets = dec_inp[1:int(seq_len.max()) + 1]
    ref_loss, dlog = om.cross_entropy_loss(ref_logits, targets, seq_len)
    ref_g, ref_denc = om.attn_decoder_bwd(dlog, cache)
    vs = VariableStore(DEV, capacity=1 << 23)
    vs.load(w)
    dec = AttnDecoder(True, model_params(cfg).decoder_params["char"], scope="char", variables=vs)
    enc_t = T(enc).requires_grad_()
    sl = T(seq_len, torch.int64)
    logits = dec(T(dec_inp, torch.int64), sl, enc_t, T(enc_len, torch.int64))
    assert relerr(logits, ref_logits) < 2e-5
    from e2e_asr_b200.losses import LossUtils
    loss = LossUtils.cross_entropy_loss(logits, T(targets, torch.int64), sl)
    assert abs(float(loss) - ref_loss) < 1e-5 * max(1, abs(ref_loss))
    loss.backward()
    assert relerr(enc_t.grad, ref_denc) < 5e-5
    for k, g in ref_g.items():
        assert relerr(vs.grad(k), g) < 5e-5, k


def test_ctc_head():
    from e2e_asr_b200.losses import LossUtils
    rng = np.random.default_rng(8)
    # S = 2 Lmax + 1 selects the sweep kernel's states-per-lane variant: 1, 2, 8, 1, 4, 12, 16, 24, 32; B > 8 spans
    # several sweep CTAs (8 utterances each)
    for (T_, B, D, C, Lmax) in [(12, 4, 6, 5, 4), (40, 3, 8, 11, 19), (150, 2, 8, 30, 70), (9, 5, 4, 3, 1),
                                (90, 11, 6, 9, 40), (400, 3, 6, 40, 180), (520, 2, 6, 50, 250),
                                (760, 2, 4, 20, 370), (1050, 2, 4, 12, 500)]:
        st = rng.standard_normal((B, T_, D))
        in_lens = rng.integers(max(2 * Lmax + 1, 2), T_ + 1, size=B)
        lab_lens = rng.integers(1, Lmax + 1, size=B)
        lab_lens[0] = Lmax
        labels = rng.integers(0, C - 1, size=(B, Lmax))
        labels[0, :2] = labels[0, 0]            # a repeat
        k = rng.standard_normal((D, C)) * 0.7
        bias = rng.standard_normal(C) * 0.3
        lg = (st @ k + bias).transpose(1, 0, 2)
        ref_lb, ref_dlg = om.ctc_loss(np.ascontiguousarray(lg), in_lens, labels, lab_lens)
        ref_dlg = ref_dlg / B
        ref_dk = np.einsum("btd,tbc->dc", st, ref_dlg)
        ref_dst = np.einsum("tbc,dc->btd", ref_dlg, k)
        Tp = T_ + 3
        buf = torch.zeros((B, Tp, D), device=DEV)
        buf[:, :T_] = T(st)
        buf.requires_grad_()
        kt, bt = T(k).requires_grad_(), T(bias).requires_grad_()
        for view in (buf[:, :T_], buf[:, :T_].transpose(0, 1)):
            for p in (buf, kt, bt):
                p.grad = None
            stash = {}
            loss = LossUtils.ctc_head_loss(view, kt, bt, T(in_lens, torch.int64), T(labels, torch.int64),
                                           T(lab_lens, torch.int64), stash)
            assert relerr(stash["loss_b"], ref_lb) < 2e-5
            assert abs(float(loss) - ref_lb.mean()) < 2e-5 * abs(ref_lb.mean())
            (loss * 2.0).backward()
            assert relerr(kt.grad, 2 * ref_dk) < 1e-4
            assert relerr(bt.grad, 2 * ref_dlg.sum((0, 1))) < 1e-4
            assert relerr(buf.grad[:, :T_], 2 * ref_dst) < 1e-4
            assert float(buf.grad[:, T_:].abs().max()) == 0


def test_prepare_input_stack_stride():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((3, 11, 5)).astype(np.float32)
    for stack, stride in [(1, 1), (3, 1), (2, 2), (1, 3)]:
        ref = om.stack_frames(x, stack)[:, ::stride]
        out = ops.prepare_input(T(x), 16, stack, stride).cpu().numpy()
        np.testing.assert_array_equal(out[:, :ref.shape[1]], ref)
        assert np.all(out[:, ref.shape[1]:] == 0)


@pytest.mark.parametrize("mode", ["tf32x3", "bf16x2"])
def test_gemm_shared_operand_split(mode):
    """e2e_split_lo / e2e_gemm_lo: one split of a tensor serves products on the tensor and on row / column views of
    it (the dX, dW_x and shifted dW_h products of a layer), with the same result as the GEMM's own pre-pass."""
    rng = np.random.default_rng(3)
    N_, I, G4 = 520, 96, 256
    x, g, w = (T(rng.standard_normal(s).astype(np.float32)) for s in ((N_, I), (N_, 2 * G4), (I, 2 * G4)))
    ops.set_gemm_mode(mode)
    try:
        x_lo, g_lo, w_lo = ops.split_lo(x), ops.split_lo(g), ops.split_lo(w)
        assert x_lo is not None and g_lo is not None
        cases = [(dict(a=g, b=w, tb=True), dict(a_lo=g_lo, b_lo=w_lo)),
                 (dict(a=x, b=g, ta=True), dict(a_lo=x_lo, b_lo=g_lo)),
                 (dict(a=x[:N_ - 1], b=g[1:, :G4], ta=True), dict(a_lo=x_lo[:N_ - 1], b_lo=g_lo[1:, :G4])),
                 (dict(a=x[1:], b=g[:N_ - 1, G4:], ta=True), dict(b_lo=g_lo[:N_ - 1, G4:]))]
        for kw, los in cases:
            plain = ops.gemm(**kw)
            shared = ops.gemm(**kw, **los)
            a64 = kw["a"].double().T if kw.get("ta") else kw["a"].double()
            b64 = kw["b"].double().T if kw.get("tb") else kw["b"].double()
            assert relerr(shared, (a64 @ b64).cpu().numpy()) < 3e-5
            assert float((shared - plain).abs().max()) <= 1e-5 * float(plain.abs().max())
    finally:
        ops.set_gemm_mode("fp32")


@pytest.mark.parametrize("K", [1024, 4096, 5000])
def test_gemm_f16x2_rows_of_very_different_magnitude(K):
    """Mode 4 scales every row of the A operand by its own power of two before the fp16 split (gradients of different
    utterances differ by many orders of magnitude) and scales the output row back: every ROW must be accurate
    relative to ITS OWN magnitude, for the GEMM's own pre-pass and for caller-held planes (e2e_split_rows_f16), and
    the shared B planes of e2e_split_lo(4) must give the same result.  K = 1024: one warp per row with the row in
    registers; 4096 / 5000: one CTA per row (the d z rows of the H = 512 encoder)."""
    rng = np.random.default_rng(11)
    M, N = 700, 512
    scale = 10.0 ** rng.uniform(-12, 2, size=(M, 1))
    a = (rng.standard_normal((M, K)) * scale).astype(np.float32)
    a[5] = 0.0
    w = (rng.standard_normal((N, K)) * 0.07).astype(np.float32)           # stored [N, K]: the dX = dz . W^T form
    ref = a.astype(np.float64) @ w.astype(np.float64).T
    row_mag = np.abs(a).max(1, keepdims=True).astype(np.float64) * np.abs(w).max() * np.sqrt(K)
    row_mag[row_mag == 0] = 1.0
    outs = [ops.gemm(T(a), T(w), tb=True, mode=4)]
    ad, wd = T(a), T(w)
    outs.append(ops.gemm(ad, wd, tb=True, mode=4, a_lo=ops.split_rows_f16(ad), b_lo=ops.split_lo(wd, 4)))
    outs.append(ops.gemm(ad, wd, tb=True, mode=4, b_lo=ops.split_lo(wd, 4)))
    for out in outs:
        err = np.abs(out.cpu().numpy().astype(np.float64) - ref) / row_mag
        assert err.max() < 2e-6, err.max()
        assert float(out[5].abs().max()) == 0.0
    if K == 1024:     # same planes, same arithmetic (larger K with this few tiles runs split-K: atomics reorder the sum)
        assert float((outs[0] - outs[1]).abs().max()) == 0.0


@pytest.mark.parametrize("mode,tol", [(1, 2e-5), (2, 2e-2), (3, 3e-5), (4, 5e-6)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 384, 96), (1000, 520, 264), (4480, 2048, 120),
                                   (256, 1024, 5000), (777, 333, 1111)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_gemm_tensor_core(mode, tol, M, N, K, ta, tb):
    """tcgen05 paths: mode 1 = 3xTF32 (fp32-accurate), mode 2 = bf16 (looser, stated tolerance), mode 3 = bf16x2
    (hi + lo bf16 pairs, three products: 16 significand bits per operand), mode 4 = f16x2 (hi + scaled-lo fp16 pairs,
    cross terms in a second accumulator: 22 significand bits; a transposed A runs mode 1)."""
    if mode == 4 and ta:
        tol = 2e-5
    rng = np.random.default_rng(M + N + K + mode)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    z = rng.standard_normal((M, N)).astype(np.float32)
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    out = ops.gemm(T(a), T(b), ta=ta, tb=tb, mode=mode)
    assert relerr(out, ref) < tol
    out = ops.gemm(T(a), T(b), ta=ta, tb=tb, bias=T(bias), z=T(z), mode=mode)
    assert relerr(out, ref + bias + z) < tol
    c0 = rng.standard_normal((M, N)).astype(np.float32)
    c = T(c0)
    ops.gemm(T(a), T(b), ta=ta, tb=tb, out=c, accumulate=True, mode=mode)
    assert relerr(c, ref + c0) < tol

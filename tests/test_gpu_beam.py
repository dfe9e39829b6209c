"""GPU beam search (float64 batched decoder step + device-side candidate merge) against the
golden ids produced by the reference's own beam_search.py (tests/golden/gen_golden.py)
and against the CPU oracle on a larger synthetic eval batch.  Token ids must be
bit-exact (north-star)."""
import os
import sys

import numpy as np
import pytest

from e2e_asr_b200 import synth
from oracle import beam as ob

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import gen_golden as gg  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gg.BEAM_CASES, ids=[c[0] for c in gg.BEAM_CASES])
def test_beam_ids_match_reference_golden(case, golden_dir):
    from e2e_asr_b200.beam_search import BeamSearch
    tag, cname, wseed, out_scale, eos_bias, n_utt, _, beams = case
    g = np.load(os.path.join(golden_dir, "beam_%s.npz" % tag))
    cfg = synth.get_config(cname)
    w = gg.dec_weights(cfg, wseed, eos_bias, out_scale)
    encs = [g["enc%d" % u] for u in range(n_utt)]
    for k in beams:
        sp = BeamSearch.class_params()
        sp.beam_size = int(k)
        bs = BeamSearch(w, sp, device="cuda:0")
        batched = bs.decode_batch(encs)                       # all utterances in one batch
        for u in range(n_utt):
            np.testing.assert_array_equal(batched[u], g["ids_k%d_u%d" % (k, u)])
        single = bs(encs[0])                                   # reference call signature, one utterance
        np.testing.assert_array_equal(single, g["ids_k%d_u0" % k])
        np.testing.assert_array_equal(bs(encs[1][None]), g["ids_k%d_u1" % k])   # [1,T,D] accepted too


def test_beam_k10_eval_batch_vs_oracle_and_lm_weight():
    from e2e_asr_b200.beam_search import BeamSearch
    cfg = synth.get_config("cfg1")
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    rng = np.random.Generator(np.random.PCG64(17))
    encs = [(np.tanh(rng.standard_normal((int(rng.integers(50, 89)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(12)]
    for k, lmw in ((10, 0.0), (4, 0.3)):
        sp = BeamSearch.class_params()
        sp.beam_size, sp.lm_weight, sp.lm_path = k, lmw, w        # the same "checkpoint" 
        out, sc = BeamSearch(w, sp, device="cuda:0").decode_batch(encs, return_scores=True)
        for u, enc in enumerate(encs):
            ref, rs = ob.beam_search(w, enc, beam_size=k, lm_weight=lmw, return_score=True)
            np.testing.assert_array_equal(out[u], ref)
            # scores agree to float32 rounding of enc.AttnW (computed in float32 by the reference too)
            assert abs(sc[u] - rs) <= 1e-6 * max(1.0, abs(rs))


@pytest.mark.parametrize("step", [0, 5])
def test_device_merge_kernel_equals_host_merge(step):
    """e2e_beam_merge (one warp per utterance, fixed hypothesis slots) against merge_candidates, the NumPy restatement
    of beam_search.py:255-266 / 294-329 that tests/test_beam_merge_cpu.py checks against the reference's per-utterance
    loop: same surviving (parent, token, score) set per utterance, same finished hypotheses, same beam sizes; scores
    bit-exact (the same float64 additions)."""
    import torch
    from e2e_asr_b200._lib import BeamMergeArgs, call
    from e2e_asr_b200.beam_search import merge_candidates
    from e2e_asr_b200.data_utils import EOS_ID
    rng = np.random.default_rng(step)
    N, beam, V, S = 37, 10, 12, 8
    R = N * beam
    k_u = rng.integers(0, beam + 1, size=N) if step else np.full(N, beam)
    alive = np.zeros(R, np.int32)
    for u in range(N):
        alive[u * beam:u * beam + (k_u[u] if step else 1)] = 1
    score = rng.standard_normal(R) * 3.0
    if step == 0:
        score[:] = 0.0
    val = np.log(rng.dirichlet(np.ones(V), size=R))
    idx = np.argsort(-val, axis=1, kind="stable")[:, :beam].astype(np.int32)          # includes EOS (= 2) now and then
    val = np.take_along_axis(val, idx.astype(np.int64), 1)
    rows = np.flatnonzero(alive)
    k_ref = k_u.astype(np.int64).copy()
    new, finished = merge_candidates(rows // beam, score[rows], k_ref, idx[rows], val[rows], step, 0.25)
    dev = "cuda:0"
    T = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    t = dict(step=T(np.array([step]), torch.int32), out_idx=T(idx, torch.int32), out_val=T(val, torch.float64),
             score=T(score, torch.float64), alive=T(alive, torch.int32), k_u=T(k_u, torch.int32),
             new_tok=torch.zeros(R, dtype=torch.int64, device=dev), new_score=torch.zeros(R, dtype=torch.float64, device=dev),
             parent=torch.zeros(R, dtype=torch.int32, device=dev), new_alive=torch.zeros(R, dtype=torch.int32, device=dev),
             krow=torch.zeros(R, dtype=torch.int32, device=dev), par_hist=torch.zeros((S, R), dtype=torch.int32, device=dev),
             tok_hist=torch.zeros((S, R), dtype=torch.int32, device=dev), fin_cnt=torch.zeros(N, dtype=torch.int32, device=dev),
             fin_step=torch.zeros(R, dtype=torch.int32, device=dev), fin_row=torch.zeros(R, dtype=torch.int32, device=dev),
             fin_score=torch.zeros(R, dtype=torch.float64, device=dev), n_live=torch.zeros(1, dtype=torch.int32, device=dev))
    a = BeamMergeArgs()
    a.N, a.beam, a.R, a.eos_id, a.word_ins_penalty = N, beam, R, EOS_ID, 0.25
    for k, v in t.items():
        setattr(a, k, v.data_ptr())
    call("e2e_beam_merge", a)
    torch.cuda.synchronize()
    h = {k: v.cpu().numpy() for k, v in t.items()}
    assert list(h["k_u"]) == list(k_ref) and int(h["n_live"][0]) == int(k_ref.sum())
    for u in range(N):
        live = [r for r in range(u * beam, (u + 1) * beam) if h["new_alive"][r]]
        assert live == list(range(u * beam, u * beam + k_ref[u]))                    # live rows first
        got = sorted((int(h["parent"][r]), int(h["new_tok"][r]), float(h["new_score"][r])) for r in live)
        sel = new["utt"] == u
        ref = sorted((int(rows[p_]), int(t_), float(s_)) for p_, t_, s_ in zip(new["parent"][sel], new["tok"][sel], new["score"][sel]))
        assert got == ref, (u, got, ref)
        assert all(h["krow"][r] == k_ref[u] for r in live)
        assert [(int(h["par_hist"][step, r]), int(h["tok_hist"][step, r])) for r in live] == \
            [(int(h["parent"][r]), int(h["new_tok"][r])) for r in live]
        fin_ref = sorted((int(rows[p_]), s_) for uu, p_, s_ in finished if uu == u)
        nf = int(h["fin_cnt"][u])
        fin_got = sorted((int(h["fin_row"][u * beam + f]), float(h["fin_score"][u * beam + f])) for f in range(nf))
        assert fin_got == fin_ref and all(h["fin_step"][u * beam + f] == step for f in range(nf))


def test_beam_separate_lm_checkpoint_and_word_insertion_penalty(tmp_path):
    """search_params.lm_path names ANOTHER checkpoint (beam_search.py:45-46: map_lm_variables(get_model_params(lm_path)))
    -- its LM-LSTM / OutputProjection / embedding drive the fusion term -- given as a dict, as an .npz path, and as a
    TF-bundle prefix; word_ins_penalty is added from step 1 on (beam_search.py:321-322), not to the step-0 scores."""
    from e2e_asr_b200.beam_search import BeamSearch
    from e2e_asr_b200.tf_checkpoint import write_checkpoint
    cfg = synth.get_config("cfg1")
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    w_lm = gg.dec_weights(cfg, 33, 1.0, 6.0)
    rng = np.random.Generator(np.random.PCG64(5))
    encs = [(np.tanh(rng.standard_normal((int(rng.integers(20, 40)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(5)]
    npz = str(tmp_path / "lm.npz")
    np.savez(npz, **w_lm)
    prefix = str(tmp_path / "lm_ckpt")
    write_checkpoint(prefix, w_lm)
    refs = [ob.beam_search(w, e, beam_size=4, lm_weight=0.4, word_ins_penalty=0.3, return_score=True, lm_weights=w_lm)
            for e in encs]
    same = [ob.beam_search(w, e, beam_size=4, lm_weight=0.4, word_ins_penalty=0.3) for e in encs]
    assert any(not np.array_equal(a[0], b) for a, b in zip(refs, same)), "the separate LM must matter in this test"
    for src in (w_lm, npz, prefix):
        sp = BeamSearch.class_params()
        sp.beam_size, sp.lm_weight, sp.lm_path, sp.word_ins_penalty = 4, 0.4, src, 0.3
        out, sc = BeamSearch(w, sp, device="cuda:0").decode_batch(encs, return_scores=True)
        for u in range(len(encs)):
            np.testing.assert_array_equal(out[u], refs[u][0])
            assert abs(sc[u] - refs[u][1]) <= 1e-6 * max(1.0, abs(refs[u][1]))
    sp = BeamSearch.class_params()
    sp.beam_size, sp.lm_weight, sp.lm_path = 4, 0.4, str(tmp_path / "missing")
    with pytest.raises(FileNotFoundError):
        BeamSearch(w, sp, device="cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,lda_pad", [(2560, 1024, 512, 0), (2560, 1000, 256, 0), (37, 64, 16, 0),
                                           (130, 257, 768, 2), (65, 129, 50, 0), (70, 40, 33, 1), (1, 8, 16, 0)])
def test_gemm_f64_matches_numpy(M, N, K, lda_pad):
    """e2e_gemm_f64 (float64 hypotheses x float32 weights + float32 bias, beam_search.py:182-199): both kernels --
    the cp.async pipeline for 16-byte-aligned operands and the plain register-tiled one -- against NumPy float64."""
    import torch
    from e2e_asr_b200._lib import call
    rng = np.random.default_rng(M * 31 + N * 7 + K)
    a = rng.standard_normal((M, K + lda_pad))
    b = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    ad = torch.from_numpy(a).cuda()[:, :K]
    bd, biasd = torch.from_numpy(b).cuda(), torch.from_numpy(bias).cuda()
    out = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
    call("e2e_gemm_f64", M, N, K, ad, ad.stride(0), bd, bd.stride(0), out, out.stride(0), biasd)
    ref = a[:, :K] @ b.astype(np.float64) + bias.astype(np.float64)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
    if K % 16 == 0 and (K + lda_pad) % 2 == 0 and N % 2 == 0:
        # the weights widened once by the caller
        out_d = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
        b64 = bd.to(torch.float64)
        from e2e_asr_b200._lib import lib
        lib().e2e_set_f64_mma(0)                 # register-tiled DFMA kernel: the same sums in the same order
        try:
            call("e2e_gemm_f64d", M, N, K, ad, ad.stride(0), b64, b64.stride(0), out_d, out_d.stride(0), biasd)
        finally:
            lib().e2e_set_f64_mma(1)
        assert torch.equal(out, out_d)
        out_t = torch.full((M, N), float("nan"), dtype=torch.float64, device="cuda")
        call("e2e_gemm_f64d", M, N, K, ad, ad.stride(0), b64, b64.stride(0), out_t, out_t.stride(0), biasd)   # DMMA
        assert np.abs(out_t.cpu().numpy() - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("beam,A,D", [(10, 128, 512), (3, 40, 72), (16, 128, 1024), (17, 32, 64)])
def test_grouped_beam_attention_is_bit_identical_to_the_per_row_kernel(beam, A, D):
    """e2e_attn_beam_group_f64 (one CTA per utterance, encoder rows read once for all its hypotheses) against
    e2e_attn_beam_f64 (one CTA per hypothesis): the same arithmetic in the same order, so exactly equal; both against
    a NumPy float64 calc_attention (beam_search.py:150-159).  beam = 17 exceeds the grouped kernel's register tile and
    must fall through to the per-row kernel."""
    import torch
    from e2e_asr_b200._lib import call
    rng = np.random.default_rng(beam * 1000 + A + D)
    Ts = np.array([44, 1, 87, 50, 9, 200, 33], np.int32)
    N, Tmax = len(Ts), int(Ts.max())
    offs = np.concatenate([[0], np.cumsum(Ts)[:-1]]).astype(np.int32)
    rows = int(Ts.sum())
    HF = rng.standard_normal((rows, A)).astype(np.float32)
    enc = np.tanh(rng.standard_normal((rows, D))).astype(np.float32)
    y = rng.standard_normal((N * beam, A))
    v = rng.standard_normal(A).astype(np.float32)
    dev = "cuda:0"
    HFd, encd, yd, vd = (torch.from_numpy(x).to(dev) for x in (HF, enc, y, v))
    ro = torch.from_numpy(np.repeat(offs, beam)).to(dev)
    rt = torch.from_numpy(np.repeat(Ts, beam)).to(dev)
    c_row = torch.full((N * beam, D), float("nan"), dtype=torch.float64, device=dev)
    c_grp = torch.full((N * beam, D), float("nan"), dtype=torch.float64, device=dev)
    call("e2e_attn_beam_f64", N * beam, A, D, Tmax, HFd, encd, ro, rt, yd, vd, c_row, D)
    call("e2e_attn_beam_group_f64", N, beam, A, D, Tmax, HFd, encd, ro, rt, yd, vd, c_grp, D)
    assert torch.equal(c_row, c_grp)
    ref = np.empty((N * beam, D))
    for r in range(N * beam):
        u = r // beam
        sl = slice(offs[u], offs[u] + Ts[u])
        s = np.tanh(HF[sl].astype(np.float64) + y[r]) @ v.astype(np.float64)
        e = np.exp(s - s.max())
        ref[r] = (e / e.sum()) @ enc[sl].astype(np.float64)
    assert np.abs(c_grp.cpu().numpy() - ref).max() <= 1e-12


@pytest.mark.gpu
def test_decode_batch_graph_capture_and_replay_repeat_the_eager_ids():
    """decode_batch keeps the buffers and the step graph of a batch signature: call 1 launches kernel by kernel, call 2
    captures the step, call 3 replays it -- the same ids and scores every time, also after a different batch of the same
    signature went through the cached buffers in between."""
    from e2e_asr_b200.beam_search import BeamSearch
    cfg = synth.get_config("cfg1")
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    rng = np.random.Generator(np.random.PCG64(5))
    lens = [int(rng.integers(50, 89)) for _ in range(6)]
    mk = lambda: [(np.tanh(rng.standard_normal((t, 2 * cfg.H))) * 0.8).astype(np.float32) for t in lens]
    encs, other = mk(), mk()
    sp = BeamSearch.class_params()
    sp.beam_size = 5
    bs = BeamSearch(w, sp, device="cuda:0")
    first, sc1 = bs.decode_batch(encs, return_scores=True)
    assert next(iter(bs._plans.values())).graph is None
    second, sc2 = bs.decode_batch(encs, return_scores=True)
    assert next(iter(bs._plans.values())).graph is not None and len(bs._plans) == 1
    o3 = bs.decode_batch(other)
    third, sc3 = bs.decode_batch(encs, return_scores=True)
    for a, b, c in zip(first, second, third):
        np.testing.assert_array_equal(a, b)
        np.testing.assert_array_equal(a, c)
    # enc . AttnW is a float32 product (beam_search.py:148) whose split-K partial sums land in arrival order: the scores
    # repeat to float32 rounding of that product, the ids exactly
    np.testing.assert_allclose(sc2, sc1, rtol=1e-6)
    np.testing.assert_allclose(sc3, sc1, rtol=1e-6)
    ref = BeamSearch(w, sp, device="cuda:0").decode_batch(other, use_graph=False)
    for a, b in zip(o3, ref):
        np.testing.assert_array_equal(a, b)
    for u, enc in enumerate(encs):
        np.testing.assert_array_equal(first[u], ob.beam_search(w, enc, beam_size=5))


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K1,K2,table", [(2560, 1024, 256, 0, True), (2560, 256, 256, 512, False),
                                              (2560, 128, 256, 0, False), (77, 1000, 256, 0, False),
                                              (2560, 1024, 256, 256, False), (33, 64, 16, 32, True)])
def test_gemm_f64d_cat_matches_numpy(M, N, K1, K2, table):
    """e2e_gemm_f64d_cat: [A1 | A2] . B + bias + Z[zrow] on the FP64 tensor cores (the reference's concatenated
    operands, basic_lstm.py:17 / beam_search.py:186,194, taken in place; 64- and 32-row tiles) against NumPy float64."""
    import torch
    from e2e_asr_b200._lib import call
    rng = np.random.default_rng(M + 3 * N + 5 * K1 + 7 * K2)
    a1 = rng.standard_normal((M, K1 + 2))
    a2 = rng.standard_normal((M, K2 + 4)) if K2 else None
    b = rng.standard_normal((K1 + K2, N)).astype(np.float32).astype(np.float64)
    bias = rng.standard_normal(N).astype(np.float32)
    z = rng.standard_normal((50, N)) if table else None
    rows = rng.integers(0, 50, M).astype(np.int64)
    dev = "cuda:0"
    a1d = torch.from_numpy(a1).to(dev)[:, :K1]
    a2d = torch.from_numpy(a2).to(dev)[:, :K2] if K2 else None
    bd, biasd = torch.from_numpy(b).to(dev), torch.from_numpy(bias).to(dev)
    zd = torch.from_numpy(z).to(dev) if table else None
    rd = torch.from_numpy(rows).to(dev)
    out = torch.full((M, N), float("nan"), dtype=torch.float64, device=dev)
    call("e2e_gemm_f64d_cat", M, N, K1, K2, a1d, a1d.stride(0), a2d, a2d.stride(0) if K2 else 0, bd, bd.stride(0), out,
         out.stride(0), biasd, zd, N if table else 0, rd if table else None)
    ref = a1[:, :K1] @ b[:K1] + bias.astype(np.float64)
    if K2:
        ref = ref + a2[:, :K2] @ b[K1:]
    if table:
        ref = ref + z[rows]
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("M,H,K1,K2,table", [(2560, 256, 256, 0, True), (2560, 256, 256, 256, False),
                                              (45, 16, 16, 16, False), (300, 64, 32, 0, True)])
def test_gemm_f64d_lstm_matches_product_plus_lstm_step(M, H, K1, K2, table):
    """e2e_gemm_f64d_lstm (BasicLSTM.__call__, basic_lstm.py:14-23, in the epilogue of the product over gate-interleaved
    columns) against e2e_gemm_f64d_cat + e2e_lstm_step_f64 in the TF column order, and against NumPy."""
    import torch
    from e2e_asr_b200._lib import call
    from e2e_asr_b200.beam_search import lstm_gate_perm
    rng = np.random.default_rng(M + H + K1 + K2)
    dev = "cuda:0"
    a1 = rng.standard_normal((M, K1)) * 0.5
    a2 = rng.standard_normal((M, K2)) * 0.5 if K2 else None
    w = (rng.standard_normal((K1 + K2, 4 * H)) * 0.2).astype(np.float32).astype(np.float64)
    bias = rng.standard_normal(4 * H).astype(np.float32)
    z = rng.standard_normal((30, 4 * H)) if table else None
    rows = rng.integers(0, 30, M).astype(np.int64)
    c_prev = rng.standard_normal((M, H))
    perm = lstm_gate_perm(H)
    t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    a1d, a2d, cd, rd = t(a1), t(a2), t(c_prev), t(rows)
    # TF column order: product, then the step kernel
    pre = torch.empty((M, 4 * H), dtype=torch.float64, device=dev)
    wd, bd, zd = t(w), t(bias), t(z)
    call("e2e_gemm_f64d_cat", M, 4 * H, K1, K2, a1d, K1, a2d, K2, wd, 4 * H, pre, 4 * H, bd, zd, 4 * H if table else 0,
         rd if table else None)
    c_ref = torch.empty((M, H), dtype=torch.float64, device=dev)
    h_ref = torch.empty((M, H), dtype=torch.float64, device=dev)
    call("e2e_lstm_step_f64", M, H, pre, cd, c_ref, h_ref, H)
    # interleaved columns: one launch
    wp, bp, zp = t(w[:, perm]), t(bias[perm]), t(None if z is None else z[:, perm])
    c_out = torch.full((M, H), float("nan"), dtype=torch.float64, device=dev)
    h_out = torch.full((M, H + 3), float("nan"), dtype=torch.float64, device=dev)
    call("e2e_gemm_f64d_lstm", M, H, K1, K2, a1d, K1, a2d, K2, wp, 4 * H, bp, zp, 4 * H if table else 0,
         rd if table else None, cd, c_out, h_out, H + 3)
    # the same sums and formulas (compiled twice: allow the last bit)
    assert (c_out - c_ref).abs().max().item() <= 1e-15 and (h_out[:, :H] - h_ref).abs().max().item() <= 1e-15
    g = a1 @ w[:K1] + bias.astype(np.float64)
    if K2:
        g = g + a2 @ w[K1:]
    if table:
        g = g + z[rows]
    sig = lambda x: 1.0 / (1.0 + np.exp(-x))
    c_np = c_prev * sig(g[:, 2 * H:3 * H] + 1.0) + sig(g[:, :H]) * np.tanh(g[:, H:2 * H])
    h_np = sig(g[:, 3 * H:]) * np.tanh(c_np)
    assert np.abs(c_out.cpu().numpy() - c_np).max() <= 1e-12 and np.abs(h_out[:, :H].cpu().numpy() - h_np).max() <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("beam,A,D,scale", [(10, 128, 512, 1.0), (3, 40, 72, 1.0), (16, 160, 1024, 1.0),
                                             (4, 128, 64, 40.0)])
def test_beam_attention_from_exponentials(beam, A, D, scale):
    """e2e_exp2x_f64 + e2e_attn_beam_group_e_f64 (tanh(h + y) = 1 - 2 / (exp(2h) exp(2y) + 1)) against the tanh kernel
    and NumPy's calc_attention (beam_search.py:150-159); scale = 40 saturates most tanh arguments."""
    import torch
    from e2e_asr_b200._lib import call
    rng = np.random.default_rng(beam * 1000 + A + D)
    Ts = np.array([44, 1, 87, 50, 9, 200, 33], np.int32)
    N, Tmax = len(Ts), int(Ts.max())
    offs = np.concatenate([[0], np.cumsum(Ts)[:-1]]).astype(np.int32)
    rows = int(Ts.sum())
    HF = (rng.standard_normal((rows, A)) * scale).astype(np.float32)
    enc = np.tanh(rng.standard_normal((rows, D))).astype(np.float32)
    y = rng.standard_normal((N * beam, A)) * scale
    v = rng.standard_normal(A).astype(np.float32)
    dev = "cuda:0"
    HFd, encd, yd, vd = (torch.from_numpy(x).to(dev) for x in (HF, enc, y, v))
    ro = torch.from_numpy(np.repeat(offs, beam)).to(dev)
    rt = torch.from_numpy(np.repeat(Ts, beam)).to(dev)
    EHF = torch.empty((rows, A), dtype=torch.float64, device=dev)
    call("e2e_exp2x_f64", rows * A, HFd, EHF)
    np.testing.assert_allclose(EHF.cpu().numpy(), np.exp(np.clip(2.0 * HF.astype(np.float64), -300.0, 300.0)), rtol=4e-16)
    c_tanh = torch.full((N * beam, D), float("nan"), dtype=torch.float64, device=dev)
    c_exp = torch.full((N * beam, D), float("nan"), dtype=torch.float64, device=dev)
    call("e2e_attn_beam_group_f64", N, beam, A, D, Tmax, HFd, encd, ro, rt, yd, vd, c_tanh, D)
    call("e2e_attn_beam_group_e_f64", N, beam, A, D, Tmax, EHF, encd, ro, rt, yd, vd, c_exp, D)
    ref = np.empty((N * beam, D))
    for r in range(N * beam):
        u = r // beam
        sl = slice(offs[u], offs[u] + Ts[u])
        s = np.tanh(HF[sl].astype(np.float64) + y[r]) @ v.astype(np.float64)
        e = np.exp(s - s.max())
        ref[r] = (e / e.sum()) @ enc[sl].astype(np.float64)
    tol = 1e-12 if scale == 1.0 else 1e-10       # saturated scores: the softmax amplifies 1e-16 of a score of ~100
    assert np.abs(c_exp.cpu().numpy() - ref).max() <= tol
    assert np.abs(c_exp.cpu().numpy() - c_tanh.cpu().numpy()).max() <= tol


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [dict(fast_step=False), dict(token_table=False, fused_lstm=False),
                                   dict(exp_attention=False, fused_lstm=False), dict(token_table=False)])
def test_beam_step_formulations_give_the_same_ids(flags):
    """The decoding step in its earlier formulations (materialised concatenations, embedding product every step, LSTM
    as its own kernel, tanh attention): the same ids as the oracle, with and without LM fusion."""
    from e2e_asr_b200.beam_search import BeamSearch
    cfg = synth.get_config("cfg1")
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    rng = np.random.Generator(np.random.PCG64(23))
    encs = [(np.tanh(rng.standard_normal((int(rng.integers(50, 89)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(5)]
    for k, lmw in ((6, 0.0), (3, 0.3)):
        sp = BeamSearch.class_params()
        sp.beam_size, sp.lm_weight, sp.lm_path = k, lmw, w
        bs = BeamSearch(w, sp, device="cuda:0")
        for name, val in flags.items():
            setattr(bs, name, val)
        out = bs.decode_batch(encs)
        for u, enc in enumerate(encs):
            np.testing.assert_array_equal(out[u], ob.beam_search(w, enc, beam_size=k, lm_weight=lmw))


@pytest.mark.gpu
@pytest.mark.parametrize("V,kmax,lm", [(17, 4, False), (1000, 10, False), (1000, 10, True), (1500, 7, True),
                                        (2500, 5, False)])
def test_logsoftmax_topk_f64(V, kmax, lm):
    """get_top_k's tail (beam_search.py:196-214): log(softmax) (+ lm_weight * log(softmax_LM)), the k best per row in
    descending order, ties to the lower index, k per row from krow (0 = dead row); register-resident kernel for
    V <= 2048, the global-memory one beyond -- against NumPy, with repeated logits to force ties."""
    import torch
    from e2e_asr_b200._lib import call
    rng = np.random.default_rng(V + kmax)
    n = 37
    x = np.round(rng.standard_normal((n, V)) * 3.0, 1)            # one decimal: many exact ties
    xl = np.round(rng.standard_normal((n, V)) * 2.0, 1)
    krow = rng.integers(0, kmax + 1, n).astype(np.int32)
    krow[0] = kmax
    dev = "cuda:0"
    xd, xld, kd = torch.from_numpy(x).to(dev), torch.from_numpy(xl).to(dev), torch.from_numpy(krow).to(dev)
    oi = torch.full((n, kmax), -7, dtype=torch.int32, device=dev)
    ov = torch.full((n, kmax), float("nan"), dtype=torch.float64, device=dev)
    scratch = torch.empty((n, V), dtype=torch.float64, device=dev)
    call("e2e_logsoftmax_topk_f64", n, V, xd, xld if lm else None, 0.3, kd, kmax, oi, ov, scratch)

    def logsm(a):
        e = np.exp(a - a.max(axis=1, keepdims=True))
        return np.log(e / e.sum(axis=1, keepdims=True))
    comb = logsm(x) + (0.3 * logsm(xl) if lm else 0.0)
    oi, ov = oi.cpu().numpy(), ov.cpu().numpy()
    for r in range(n):
        k = int(krow[r])
        assert (oi[r, k:] == -1).all() and np.isneginf(ov[r, k:]).all()
        got = oi[r, :k]
        assert len(set(got.tolist())) == k
        np.testing.assert_allclose(ov[r, :k], comb[r, got], rtol=0, atol=1e-12)
        # descending values; equal device values in ascending index order; nothing left out that beats the k-th
        for a, b in zip(range(k - 1), range(1, k)):
            assert ov[r, a] > ov[r, b] or (ov[r, a] == ov[r, b] and got[a] < got[b])
        if 0 < k < V:
            rest = np.delete(comb[r], got)
            assert rest.max() <= ov[r, k - 1] + 1e-12


@pytest.mark.gpu
def test_device_resident_encoder_states_give_the_same_ids():
    """Encoder states passed as CUDA tensors (an evaluation loop feeding the encoder's output) are gathered on the
    device; [1, T, D] accepted; same ids as from host arrays."""
    import torch
    from e2e_asr_b200.beam_search import BeamSearch
    cfg = synth.get_config("cfg1")
    w = gg.dec_weights(cfg, 21, 2.5, 10.0)
    rng = np.random.Generator(np.random.PCG64(31))
    encs = [(np.tanh(rng.standard_normal((int(rng.integers(50, 89)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(4)]
    sp = BeamSearch.class_params()
    sp.beam_size = 4
    bs = BeamSearch(w, sp, device="cuda:0")
    host = bs.decode_batch(encs)
    devt = [torch.from_numpy(e).cuda() for e in encs]
    devt[1] = devt[1][None]
    on_dev = bs.decode_batch(devt)
    for a, b in zip(host, on_dev):
        np.testing.assert_array_equal(a, b)

"""Oracle vs golden vectors produced by executing the reference's own NumPy code
(tests/golden/gen_golden.py: basic_lstm.py, num_utils.py, beam_search.py)."""
import os
import sys

import numpy as np
import pytest

from e2e_asr_b200 import synth
from oracle import beam as ob
from oracle import model as om

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import gen_golden as gg  # noqa: E402  (only dec_weights/BEAM_CASES; never touches /root/reference here)


def test_lstm_cell_and_numerics(golden_dir):
    g = np.load(os.path.join(golden_dir, "cell_numerics.npz"))
    for i in range(3):
        nc, nh = om.lstm_cell(g["lstm%d_x" % i][None], g["lstm%d_c" % i][None], g["lstm%d_h" % i][None],
                              g["lstm%d_w" % i], g["lstm%d_b" % i])
        np.testing.assert_allclose(nc[0], g["lstm%d_nc" % i], rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(nh[0], g["lstm%d_nh" % i], rtol=1e-13, atol=1e-15)
        nc2, nh2 = ob.basic_lstm(g["lstm%d_x" % i], (g["lstm%d_c" % i], g["lstm%d_h" % i]),
                                 g["lstm%d_w" % i], g["lstm%d_b" % i])
        np.testing.assert_array_equal(nc2, g["lstm%d_nc" % i])
        np.testing.assert_array_equal(nh2, g["lstm%d_nh" % i])
    np.testing.assert_array_equal(om.sigmoid(g["num_x"]), g["num_sigmoid"])
    np.testing.assert_array_equal(om.softmax(g["num_x"]), g["num_softmax"])


@pytest.mark.parametrize("case", gg.BEAM_CASES, ids=[c[0] for c in gg.BEAM_CASES])
def test_beam_search_matches_reference(case, golden_dir):
    tag, cname, wseed, out_scale, eos_bias, n_utt, _, beams = case
    g = np.load(os.path.join(golden_dir, "beam_%s.npz" % tag))
    cfg = synth.get_config(cname)
    w = gg.dec_weights(cfg, wseed, eos_bias, out_scale)
    lens = set()
    for k in beams:
        for u in range(n_utt):
            ids = ob.beam_search(w, g["enc%d" % u], beam_size=int(k))
            np.testing.assert_array_equal(ids, g["ids_k%d_u%d" % (k, u)])
            lens.add(len(ids))
    assert len(lens) >= 2          # fixtures cover both early-EOS and long outputs


@pytest.mark.parametrize("case", gg.BEAM_CASES, ids=[c[0] for c in gg.BEAM_CASES])
def test_decoder_step_matches_reference(case, golden_dir):
    """One get_top_k call from GO and a second from its first result
    (beam_search.py:178-219) incl. the lm_weight term."""
    tag, cname, wseed, out_scale, eos_bias, n_utt, _, beams = case
    g = np.load(os.path.join(golden_dir, "beam_%s.npz" % tag))
    cfg = synth.get_config(cname)
    w = gg.dec_weights(cfg, wseed, eos_bias, out_scale)
    p = ob.DecParams(w)
    fn = ob.make_step_fn(p, g["enc0"], lm_weight=0.3)
    hs = p.dec_lstm_w.shape[1] // 4
    ls = p.lm_lstm_w.shape[1] // 4
    z = lambda n: (np.zeros(n), np.zeros(n))
    x = p.embedding[1]
    top, ms, _, st, ctx, _ = fn(x, x, [z(hs), z(ls), z(ls)], np.zeros(g["enc0"].shape[1]), 3)
    np.testing.assert_array_equal(np.sort(top), g["step_top"])
    np.testing.assert_allclose(np.sort(ms), g["step_scores"], rtol=1e-12)
    np.testing.assert_allclose(ctx, g["step_ctx"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(st[0][0], g["step_dec_c"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(st[0][1], g["step_dec_h"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(st[1][1], g["step_lm_h"], rtol=1e-12, atol=1e-14)
    x2 = p.embedding[int(g["step2_in"])]
    top2, ms2, _, st2, ctx2, _ = fn(x2, x2, st, ctx, 3)
    np.testing.assert_array_equal(np.sort(top2), g["step2_top"])
    np.testing.assert_allclose(np.sort(ms2), g["step2_scores"], rtol=1e-12)
    np.testing.assert_allclose(ctx2, g["step2_ctx"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(st2[0][0], g["step2_dec_c"], rtol=1e-12, atol=1e-14)


def test_training_decoder_step_equals_beam_step():
    """The TF-graph restatement (oracle.model.attn_decoder_fwd, unpinned) and the
    reference-pinned NumPy decoder step must agree: greedy decode through the
    training-graph decoder == beam search with k=1."""
    for cname, seed in (("tiny", 11), ("tiny_b", 12)):
        cfg = synth.get_config(cname)
        w = gg.dec_weights(cfg, seed, 2.0, 14.0)
        rng = np.random.Generator(np.random.PCG64(3))
        for _ in range(6):
            T = int(rng.integers(3, 15))
            enc = (np.tanh(rng.standard_normal((T, 2 * cfg.H))) * 0.8).astype(np.float32)
            ids = ob.beam_search(w, enc, beam_size=1)
            W64 = {k: v.astype(np.float64) for k, v in w.items()}
            dec_inp = np.full((121, 1), 1, np.int64)
            logits, _ = om.attn_decoder_fwd(W64, "char", dec_inp, np.array([120]), enc[None].astype(np.float64),
                                            np.array([T]), mode="greedy")
            g_ids = ob.greedy_ids_from_logits(logits, 1)[0]
            cut = ob.cut_at_eos(g_ids)
            # beam ids include the trailing EOS (beam_search.py:338); greedy ids are cut before it
            ref = list(ids[:-1]) if ids[-1] == ob.EOS_ID else list(ids)
            assert cut[:len(ref)] == ref
            if ids[-1] == ob.EOS_ID:
                assert len(cut) == len(ref)

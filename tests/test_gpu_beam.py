"""GPU beam search (float64 batched decoder step + host candidate merge) against the
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
        sp.beam_size, sp.lm_weight, sp.lm_path = k, lmw, "same-checkpoint"
        out, sc = BeamSearch(w, sp, device="cuda:0").decode_batch(encs, return_scores=True)
        for u, enc in enumerate(encs):
            ref, rs = ob.beam_search(w, enc, beam_size=k, lm_weight=lmw, return_score=True)
            np.testing.assert_array_equal(out[u], ref)
            # scores agree to float32 rounding of enc.AttnW (computed in float32 by the reference too)
            assert abs(sc[u] - rs) <= 1e-6 * max(1.0, abs(rs))

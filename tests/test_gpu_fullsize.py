"""Parity at BASELINE.json's STATED sizes against the float64 oracle (tests/golden/gen_fullsize_golden.py ran it once in
the build container and stored losses, the gradient norm, per-variable summaries of the clipped gradient -- max-abs, L2,
random-sign projections, a strided sample -- and samples of the logits and of the top encoder states):

  cfg-2  configs[1]  B=64, T=700, H=256, L=4 (full size; eager step AND its CUDA-graph replay)
  cfg-4  configs[3]  8 utterances at full T=2000
  cfg-5  configs[4]  16 utterances at full H=512 / L=5 / T=700 (the H=512 recurrence and D=1024 decoder kernels)
  cfg-3  configs[2]  beam search k=10, ALL 256 utterances of the eval batch, ids bit-exact

Tolerance (north-star): fp32-accurate mode (3xTF32) within 1e-4 relative -- max-abs error over max-abs value per tensor;
variables whose whole gradient is below 1e-4 of the model's largest gradient entry are measured against that floor.
"""
import os
import sys

import numpy as np
import pytest

from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import gen_fullsize_golden as fg  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def check_against_golden(model, g, cfg, rtol=RTOL):
    import torch
    torch.cuda.synchronize()
    ops.check_device_errors(model.device)
    for t, l in model.losses.items():
        ref = float(g["loss/" + t])
        assert abs(float(l.detach()) - ref) <= rtol * max(1.0, abs(ref)), ("loss", t, float(l.detach()), ref)
    ref = float(g["total_loss"])
    assert abs(float(model.total_loss) - ref) <= rtol * max(1.0, abs(ref)), ("total_loss", float(model.total_loss), ref)
    ref = float(g["norm"])
    assert abs(float(model.grad_norm) - ref) <= rtol * max(1.0, ref), ("norm", float(model.grad_norm), ref)
    names = [str(n) for n in g["names"]]
    grads = model.gradients()
    assert sorted(grads) == names
    gmax = max(float(g["maxabs/" + k]) for k in names)
    worst = 0.0
    for k in names:
        got = np.asarray(grads[k], np.float64).ravel()
        floor = max(float(g["maxabs/" + k]), 1e-4 * gmax)
        idx = fg.sample_index(got.size)
        e = float(np.abs(got[idx] - g["sample/" + k].astype(np.float64)).max()) / floor
        # the stored sample is float32: allow its own rounding on top of the tolerance
        assert e <= rtol + 2e-7, ("grad sample", k, e)
        worst = max(worst, e)
        l2 = float(np.sqrt((got * got).sum()))
        assert abs(l2 - float(g["l2/" + k])) <= rtol * max(float(g["l2/" + k]), floor), ("grad l2", k, l2)
        proj = fg.proj_signs(k, got.size) @ got
        tol = 4.0 * rtol * floor * np.sqrt(got.size)
        assert float(np.abs(proj - g["proj/" + k]).max()) <= tol, ("grad projection", k)
    lg = model.outputs["char"].detach().cpu().numpy()
    e = float(np.abs(lg[g["logit_rows"]].astype(np.float64) - g["logits"]).max()) / float(g["logits_maxabs"])
    assert e <= rtol + 2e-7, ("logits", e)
    st = model.encoder_hidden_states[cfg.L].detach().cpu().numpy().astype(np.float64).ravel()
    e = float(np.abs(st[g["states_idx"]] - g["states"]).max()) / float(g["states_maxabs"])
    assert e <= rtol + 2e-7, ("encoder states", e)
    return worst


@pytest.mark.parametrize("gemm", ["tf32x3", "f16x2"])
@pytest.mark.parametrize("tag", ["cfg2", "cfg4", "cfg5"])
def test_fullsize_step_matches_oracle_golden(tag, gemm, golden_dir):
    """Both fp32-accurate tensor-core modes: 3xTF32 and f16x2 (fp16 hi + scaled-lo pairs at half the tensor-pipe cost;
    the weight-gradient products run 3xTF32 there too)."""
    g = np.load(os.path.join(golden_dir, "fullsize_%s.npz" % tag))
    cfg = fg.case_config(tag)
    w = synth.make_weights(cfg)
    batch = synth.make_batch(cfg)
    ops.set_gemm_mode(gemm)
    try:
        model = build_model(cfg, w, device="cuda:0")
        model.run_step(batch)
        check_against_golden(model, g, cfg)
        if tag == "cfg2":           # the benchmarked form of the step: replayed from its CUDA graph
            step = model.graphed_step(batch)
            step(batch)
            check_against_golden(model, g, cfg)
    finally:
        ops.set_gemm_mode("fp32")


def test_fullsize_beam_ids_all_256_utterances(golden_dir):
    """BASELINE configs[2]: beam width 10, the bench's synthetic eval batch of 256 utterances; ids bit-exact against the
    oracle restatement of beam_search.py (itself pinned to the executed reference by beam_*.npz)."""
    from e2e_asr_b200.beam_search import BeamSearch
    g = np.load(os.path.join(golden_dir, "fullsize_beam.npz"))
    cfg = synth.get_config("cfg2")
    w = synth.make_weights(cfg)
    encs = synth.make_beam_eval_batch(cfg, fg.BEAM_UTTS)
    assert [e.shape[0] for e in encs] == list(g["enc_lens"])
    sp = BeamSearch.class_params()
    sp.beam_size = int(g["beam_size"])
    out = BeamSearch(w, sp, device="cuda:0").decode_batch(encs)
    off = np.concatenate([[0], np.cumsum(g["lens"])])
    bad = [u for u in range(len(encs)) if not np.array_equal(out[u], g["ids"][off[u]:off[u + 1]])]
    assert not bad, "beam ids differ from the oracle for utterances %s" % bad[:10]

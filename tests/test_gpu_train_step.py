"""Whole-step parity on the GPU: losses, logits and clipped gradients of one
teacher-forced fwd+bwd step against the CPU oracle (north-star: fp32 within 1e-4
relative)."""
import numpy as np
import pytest
import torch

from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model, compare_step
from oracle import model as om

pytestmark = pytest.mark.gpu
RTOL = 1e-4    # north-star tolerance for fp32 losses / logits / gradients


@pytest.mark.parametrize("cname,ctc", [("tiny", True), ("tiny_b", True), ("tiny", False), ("cfg1", True)])
def test_train_step_matches_oracle(cname, ctc):
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0", ctc=ctc)
    model.run_step(batch)
    ops.check_device_errors("cuda:0")
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc if ctc else {})
    worst = compare_step(model, ref, rtol=RTOL)
    # reference-shaped public tensors
    d = cfg.L
    assert tuple(model.encoder_hidden_states[d].shape) == ref["states"][d].shape
    assert tuple(model.outputs["char"].shape) == ref["logits"]["char"].shape
    # second step on the same batch is bit-identical in the loss (buffers are re-zeroed)
    l1 = float(model.total_loss)
    model.run_step(batch)
    assert abs(float(model.total_loss) - l1) <= 1e-6 * abs(l1)
    print(cname, "worst grad rel err", worst)


def test_clipping_active_and_dense_norm_option():
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": 4}, ctc_tasks=cfg.ctc, max_gradient_norm=0.5)
    model = build_model(cfg, w, device="cuda:0")
    model.params.max_gradient_norm = 0.5
    model.run_step(batch)
    assert ref["norm"] > 0.5
    compare_step(model, ref, rtol=RTOL)
    model.params.tf_indexed_slices_norm = False
    model.run_step(batch)
    assert abs(float(model.grad_norm) - ref["dense_norm"]) < 1e-4 * ref["dense_norm"]

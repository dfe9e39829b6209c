"""Parity at BASELINE.json's FULL sizes (cfg-2: B=64, T=700; cfg-4: B=32, T=2000), where the float64 oracle is too slow
to run: size-independent properties of the training step that the reference graph has by construction.

* utterance-permutation equivariance: the loss is a batch mean of per-utterance terms (losses.py:32-35) and every
  kernel treats rows independently, so permuting the utterances of a batch must leave total_loss, the gradient norm and
  every clipped gradient unchanged (up to reduction order);
* padding invariance: frames past logmel_len are masked everywhere (dynamic_rnn sequence_length, attn_mask), so
  appending all-zero frames to the padded batch must not change anything;
* repeatability: the same batch twice agrees to fp32 round-off (the split-K GEMMs accumulate with fp32 atomics, so
  the summation order -- and only that -- varies between runs).
The small-size twins of these runs are checked element-wise against the oracle in test_gpu_train_step.py."""
import numpy as np
import pytest
import torch

from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model

pytestmark = pytest.mark.gpu


def _step(model, batch):
    model.run_step(batch)
    ops.check_device_errors("cuda:0")
    g = model.variables.flat_grads().detach().clone()
    return float(model.total_loss), float(model.grad_norm), g, {t: float(l) for t, l in model.losses.items()}


def _permute(batch, perm):
    return {k: (v[perm] if hasattr(v, "shape") and len(v) == len(perm) else v) for k, v in batch.items()}


@pytest.mark.parametrize("cname,gemm", [("cfg2", "tf32x3"), ("cfg4", "tf32x3")])
def test_full_size_permutation_padding_determinism(cname, gemm):
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg)
    batch = synth.make_batch(cfg)
    ops.set_gemm_mode(gemm)
    try:
        model = build_model(cfg, w, device="cuda:0")
        l0, n0, g0, losses0 = _step(model, batch)
        assert np.isfinite(l0) and n0 > 0
        gmax = float(g0.abs().max())
        # repeatability
        l0b, n0b, g0b, _ = _step(model, batch)
        assert abs(l0b - l0) <= 1e-6 * abs(l0) and abs(n0b - n0) <= 1e-5 * n0
        assert float((g0b - g0).abs().max()) <= 1e-5 * gmax
        # permutation of the utterances
        perm = np.random.Generator(np.random.PCG64(3)).permutation(cfg.B)
        l1, n1, g1, losses1 = _step(model, _permute(batch, perm))
        assert abs(l1 - l0) <= 2e-6 * abs(l0), (l0, l1)
        for t in losses0:
            assert abs(losses1[t] - losses0[t]) <= 2e-6 * max(1.0, abs(losses0[t]))
        assert abs(n1 - n0) <= 1e-5 * n0
        assert float((g1 - g0).abs().max()) <= 1e-4 * gmax
        # zero padding appended to the time axis (lengths unchanged)
        padded = dict(batch)
        x = batch["logmel"]
        padded["logmel"] = np.concatenate([x, np.zeros((x.shape[0], 8, x.shape[2]), x.dtype)], axis=1)
        l2, n2, g2, _ = _step(model, padded)
        assert abs(l2 - l0) <= 2e-6 * abs(l0), (l0, l2)
        assert abs(n2 - n0) <= 1e-5 * n0
        assert float((g2 - g0).abs().max()) <= 1e-4 * gmax
    finally:
        ops.set_gemm_mode("fp32")

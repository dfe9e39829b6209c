"""Full-size accuracy of a GEMM mode: one cfg step in `mode` against the same step in tf32x3 (fp32-accurate)."""
import sys
import numpy as np
import torch
from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model

cname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
modes = sys.argv[2:] or ["bf16x2"]
cfg = synth.get_config(cname)
w = synth.make_weights(cfg, bias_noise=0.1)
batch = synth.make_batch(cfg)
res = {}
for mode in ["tf32x3"] + modes:
    ops.set_gemm_mode(mode)
    model = build_model(cfg, w, device="cuda:0")
    model.run_step(batch)
    torch.cuda.synchronize()
    res[mode] = (float(model.total_loss), float(model.grad_norm), model.gradients(),
                 {t: model.outputs[t].detach().cpu().numpy() for t in model.params.tasks})
    del model
ref = res["tf32x3"]
for mode in modes:
    r = res[mode]
    gmax = max(float(np.abs(g).max()) for g in ref[2].values())
    worst, wk = 0.0, None
    for k, g in ref[2].items():
        e = float(np.abs(r[2][k].astype(np.float64) - g).max()) / max(float(np.abs(g).max()), 1e-4 * gmax)
        if e > worst:
            worst, wk = e, k
    le = max(float(np.abs(r[3][t] - ref[3][t]).max() / np.abs(ref[3][t]).max()) for t in ref[3])
    print(cname, mode, "loss", r[0], ref[0], "rel", abs(r[0] - ref[0]) / abs(ref[0]), "norm rel",
          abs(r[1] - ref[1]) / ref[1], "worst grad", worst, wk, "logits", le)

import sys
import numpy as np, torch
from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model
cfg = synth.get_config(sys.argv[1] if len(sys.argv) > 1 else "tiny_b")
w = synth.make_weights(cfg, bias_noise=0.1)
batch = synth.make_batch(cfg)
model = build_model(cfg, w, device="cuda:0")
try:
    gs = model.graphed_step(batch)
    gs.step(); torch.cuda.synchronize()
    print("graph ok", gs.launches_per_step, float(model.total_loss))
except Exception as e:
    print("FAILED:", str(e)[:300])

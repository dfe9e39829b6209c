import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model
ops.set_gemm_mode("tf32x3")
cfg = synth.get_config("cfg1", B=8, T=120)
w = synth.make_weights(cfg)
base = synth.make_batch(cfg, seed=100)
batches = []
for i in range(4):
    b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in base.items()}
    b["logmel"] = (base["logmel"] * (1.0 + 0.25 * i)).astype(np.float32)
    tok = b["char"] >= 3
    b["char"][tok] = (b["char"][tok] - 3 + 7 * i) % (cfg.V - 3) + 3
    batches.append(b)
model = build_model(cfg, w, device="cuda:0")
def show(tag):
    torch.cuda.synchronize()
    print(tag, {t: float(l.detach()) for t, l in model.losses.items()}, float(model.total_loss), float(model.grad_norm))
for i, b in enumerate(batches):
    model.run_step(b); show("eager b%d" % i)
step = model.graphed_step(batches[0])
show("after capture")
for rep in range(2):
    for i, b in enumerate(batches):
        step(b); show("graph b%d" % i)
for i, b in enumerate(batches):
    model.run_step(b); show("eager b%d" % i)

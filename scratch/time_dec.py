import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import _lib, ops, synth
from e2e_asr_b200.testing import build_model
ops.set_gemm_mode("tf32x3")
cfg = synth.get_config("cfg2")
model = build_model(cfg, synth.make_weights(cfg), device="cuda:0")
batch = synth.make_batch(cfg)
prepared = model.get_batch(batch)
for _ in range(2): model.run_step(prepared=prepared)
torch.cuda.synchronize()
dbg = torch.zeros(32 * 200, dtype=torch.int64, device="cuda:0")
_lib.lib().e2e_set_rec_debug(dbg.data_ptr())
# the decoder bwd overwrites the fwd stamps: run fwd only by reading after a step => bwd stamps; use hooks
import e2e_asr_b200.ops as O
orig = O.call
snap = {}
def call2(name, *a, **k):
    r = orig(name, *a, **k)
    if name.startswith("e2e_decoder_persist"):
        torch.cuda.synchronize(); snap[name] = dbg.cpu().numpy().copy()
    return r
O.call = call2
model.run_step(prepared=prepared)
torch.cuda.synchronize()
_lib.lib().e2e_set_rec_debug(0)
for name, d in snap.items():
    d = d.reshape(-1, 32)[:cfg.U]
    x = d[5:-5]
    seg = [np.median(x[:, i + 1] - x[:, i]) for i in range(6)]
    print(name, "per-step cycles: phase1 %.0f | bar %.0f | phase2 %.0f | bar %.0f | phase3 %.0f | bar %.0f | total %.0f"
          % (*seg, np.median(np.abs(np.diff(d[5:-5, 0])))))
d = snap["e2e_decoder_persist_fwd"].reshape(-1, 32)[5:cfg.U - 5]
print("fwd phase A detail: issue %.0f | wait HF %.0f | scores %.0f | softmax %.0f | issue2+wait enc0 %.0f | ctx0 %.0f | wait enc1 %.0f | ctx1+store %.0f" % tuple(
    np.median(d[:, j] - d[:, i]) for i, j in [(4, 8), (8, 9), (9, 10), (10, 11), (11, 12), (12, 13), (13, 14), (14, 5)]))
print("fwd phase G detail: issue A-tile %.0f | wait %.0f | mma %.0f | sync %.0f | epilogue %.0f" % tuple(
    np.median(d[:, j] - d[:, i]) for i, j in [(0, 16), (16, 17), (17, 18), (18, 19), (19, 1)]))

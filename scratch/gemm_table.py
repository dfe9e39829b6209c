import sys; sys.path.insert(0, '.')
import torch
from e2e_asr_b200 import _lib, ops, synth
from e2e_asr_b200.testing import build_model
ops.set_gemm_mode(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
ops.TAG_GEMM_SHAPES = True
cfg = synth.get_config("cfg2")
model = build_model(cfg, synth.make_weights(cfg), device="cuda:0")
batch = synth.make_batch(cfg)
prepared = model.get_batch(batch)
for _ in range(3): model.run_step(prepared=prepared)
torch.cuda.synchronize()
prof = _lib.Profiler(); _lib.PROFILER = prof
K = 3
for _ in range(K): model.run_step(prepared=prepared)
summ = prof.summary(); _lib.PROFILER = None
tot = 0
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    if k.startswith("gemm"):
        tot += v["ms"] / K
        print("%-40s calls %4.1f  ms/step %7.3f  TF/s %7.1f" % (k, v["calls"] / K, v["ms"] / K, v["work"] / v["ms"] / 1e9))
print("total gemm ms/step", tot)

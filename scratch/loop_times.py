import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import _lib, ops, synth
from e2e_asr_b200.testing import build_model
ops.set_gemm_mode("tf32x3")
cfg = synth.get_config("cfg2")
model = build_model(cfg, synth.make_weights(cfg), device="cuda:0")
batch = synth.make_batch(cfg)
prepared = model.get_batch(batch)
for _ in range(3): model.run_step(prepared=prepared)
torch.cuda.synchronize()
def loop(K, sync):
    st0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(K):
        model.run_step(prepared=prepared)
        if sync: torch.cuda.synchronize()
    e1.record(); th = time.perf_counter() - t0
    torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats()
    print("sync=%d: %.2f ms/step (host enqueue %.2f ms/step), cudaMalloc calls %d, reserved %.1f GB, alloc_retries %d" % (
        sync, e0.elapsed_time(e1) / K, th * 1e3 / K, st1["num_device_alloc"] - st0["num_device_alloc"],
        st1["reserved_bytes.all.current"] / 2**30, st1["num_alloc_retries"] - st0["num_alloc_retries"]))
loop(10, True); loop(10, False); loop(10, False); loop(10, True)

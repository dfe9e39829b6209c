import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import synth
from e2e_asr_b200.beam_search import BeamSearch
cfg = synth.get_config("cfg2")
w = synth.make_weights(cfg)
rng = np.random.Generator(np.random.PCG64(17))
encs = [(np.tanh(rng.standard_normal((int(rng.integers(50, 89)), 2 * cfg.H))) * 0.8).astype(np.float32) for _ in range(256)]
sp = BeamSearch.class_params(); sp.beam_size = 10
bs = BeamSearch(w, sp, device="cuda:0")
bs.decode_batch(encs[:8])
torch.cuda.synchronize()
for n in (32, 256):
    t0 = time.perf_counter(); out = bs.decode_batch(encs[:n]); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("n=%d: %.2f s, %.1f utt/s, mean len %.1f" % (n, dt, n / dt, np.mean([len(o) for o in out])))
from oracle import beam as ob
t0 = time.perf_counter(); ref = ob.beam_search(w, encs[0], beam_size=10); dt = time.perf_counter() - t0
print("oracle 1 utt: %.2f s, len %d, match %s" % (dt, len(ref), np.array_equal(ref, out[0])))
from e2e_asr_b200 import _lib
prof = _lib.Profiler(); _lib.PROFILER = prof
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = bs.decode_batch(encs[:256]); e1.record(); torch.cuda.synchronize()
_lib.PROFILER = None
summ = prof.summary()
print("total ms", e0.elapsed_time(e1))
for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"]):
    print("%-28s %8.2f ms %6d calls" % (k, v["ms"], v["calls"]))

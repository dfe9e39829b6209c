"""Per-kernel event times of one beam-search decode (eager, no graph) -- where a decoding step's time goes."""
import sys; sys.path.insert(0, '.')
import time, numpy as np, torch
from e2e_asr_b200 import _lib, synth
from e2e_asr_b200.beam_search import BeamSearch
cfg = synth.get_config("cfg2")
w = synth.make_weights(cfg)
encs = synth.make_beam_eval_batch(cfg, 256)
sp = BeamSearch.class_params(); sp.beam_size = 10
bs = BeamSearch(w, sp, device="cuda:0")
bs.decode_batch(encs[:8])
for graph in (True, False):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bs.decode_batch(encs, use_graph=graph)
    torch.cuda.synchronize(); print("graph" if graph else "eager", "%.1f ms" % ((time.perf_counter() - t0) * 1e3))
prof = _lib.Profiler(); _lib.PROFILER = prof
bs.decode_batch(encs, use_graph=False)
_lib.PROFILER = None
for k, v in sorted(prof.summary().items(), key=lambda kv: -kv[1]["ms"]):
    print("%-28s %8.2f ms  %5d calls  %.3f ms/call" % (k, v["ms"], v["calls"], v["ms"] / v["calls"]))

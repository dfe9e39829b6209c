import sys; sys.path.insert(0, '.')
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
# main-stream timeline: record an event after every C call on the current (main) stream
import e2e_asr_b200._lib as L
orig = L.call
marks = []
def call2(name, *a, **k):
    r = orig(name, *a, **k)
    if torch.cuda.current_stream() == torch.cuda.default_stream():
        e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((k.get("tag") or name, e))
    return r
for mod in (L, ops):
    mod.call = call2
import e2e_asr_b200.seq2seq_model as S, e2e_asr_b200.attn_decoder as AD, e2e_asr_b200.encoder as EN
for mod in (S, AD, EN):
    if hasattr(mod, "call"): mod.call = call2
e0 = torch.cuda.Event(enable_timing=True); e0.record()
model.run_step(prepared=prepared)
e1 = torch.cuda.Event(enable_timing=True); e1.record()
torch.cuda.synchronize()
print("step %.3f ms, %d main-stream calls" % (e0.elapsed_time(e1), len(marks)))
prev = e0; tprev = 0.0
rows = []
for name, e in marks:
    tt = e0.elapsed_time(e)
    rows.append((name, tt - tprev, tt)); tprev = tt
# print big gaps / durations
for name, d, tt in rows:
    if d > 0.15: print("%8.3f  +%.3f  %s" % (tt, d, name))

import sys, os; sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import numpy as np, torch
from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model
import gen_fullsize_golden as fg
which = sys.argv[1:]
def run(tag, graph):
    g = np.load("tests/golden/fullsize_%s.npz" % tag)
    cfg = fg.case_config(tag)
    model = build_model(cfg, synth.make_weights(cfg), device="cuda:0")
    batch = synth.make_batch(cfg)
    model.run_step(batch)
    if graph:
        st = model.graphed_step(batch); st(batch)
    torch.cuda.synchronize()
    print(tag, "norm", float(model.grad_norm), float(g["norm"]))
ops.set_gemm_mode("tf32x3")
for w_ in which:
    run(w_.rstrip("g"), w_.endswith("g"))
import gc; gc.collect()
g = np.load("tests/golden/fullsize_cfg2.npz")
cfg = synth.get_config("cfg2")
model = build_model(cfg, synth.make_weights(cfg), device="cuda:0")
batch = synth.make_batch(cfg)
for rep in range(2):
    model.run_step(batch); torch.cuda.synchronize()
    print("fresh cfg2 run", rep, "norm", float(model.grad_norm), "golden", float(g["norm"]))
    grads = model.gradients()
    for k in sorted(grads):
        l2 = float(np.sqrt((grads[k].astype(np.float64) ** 2).sum())); ref = float(g["l2/" + k])
        if abs(l2 - ref) > 1e-3 * max(ref, 1e-3): print("   BAD", k, l2, ref)

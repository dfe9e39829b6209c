import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops, _lib
DEV='cuda:0'
B,T_,I,H = 64,350,1024,256
rng = np.random.default_rng(0)
lens = rng.integers(int(0.6*T_), T_+1, size=B); lens[0]=T_
Tp = T_+2
xp = torch.zeros((B,Tp,I), device=DEV); xp[:, :T_] = torch.randn((B,T_,I), device=DEV)
xp.requires_grad_(True)
ps = [torch.empty((I+H,4*H), device=DEV).uniform_(-0.075,0.075).requires_grad_(), torch.zeros(4*H, device=DEV).requires_grad_(),
      torch.empty((I+H,4*H), device=DEV).uniform_(-0.075,0.075).requires_grad_(), torch.zeros(4*H, device=DEV).requires_grad_()]
ops.set_gemm_mode("tf32x3")
lens_t = torch.tensor(lens, dtype=torch.int32, device=DEV)
outs = {}
for mode in [2, 4, 5, 6, 0]:
    _lib.lib().e2e_set_rec_mode(mode)
    for it in range(3):
        prof = _lib.Profiler(); _lib.PROFILER = prof
        out = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max()))
        out.backward(torch.ones_like(out))
        summ = prof.summary(); _lib.PROFILER = None
    outs[mode] = (out.detach().clone(), xp.grad.detach().clone()); xp.grad = None
    print("mode %d: rec fwd %.3f ms (%.2f us/step), rec bwd %.3f ms (%.2f us/step)" % (
        mode, summ["enc_rec_fwd"]["ms"], summ["enc_rec_fwd"]["ms"] * 1e3 / T_, summ["enc_rec_bwd"]["ms"], summ["enc_rec_bwd"]["ms"] * 1e3 / T_))
for m in [4, 5, 6, 0]:
    print("mode", m, "vs 2: out maxdiff %.3e  dx maxdiff %.3e" % ((outs[m][0]-outs[2][0]).abs().max().item(), (outs[m][1]-outs[2][1]).abs().max().item()))

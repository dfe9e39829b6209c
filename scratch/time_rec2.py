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
for mode in [2, 3, 4, 0]:
    _lib.lib().e2e_set_rec_mode(mode)
    ops.reset_timing() if hasattr(ops, "reset_timing") else None
    for it in range(3):
        out = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max()))
        g = torch.ones_like(out)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out.backward(g); e1.record(); torch.cuda.synchronize()
        tb = e0.elapsed_time(e1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out2 = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max())); e1.record(); torch.cuda.synchronize()
        tf = e0.elapsed_time(e1)
    outs[mode] = (out.detach().clone(), xp.grad.detach().clone()); xp.grad = None
    print("mode %d: layer fwd %.3f ms, layer bwd %.3f ms  (T=%d)" % (mode, tf, tb, T_))
for m in [3, 4, 0]:
    print("mode", m, "vs 2: out maxdiff %.3e  dx maxdiff %.3e" % ((outs[m][0]-outs[2][0]).abs().max().item(), (outs[m][1]-outs[2][1]).abs().max().item()))

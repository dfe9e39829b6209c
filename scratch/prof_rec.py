import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops
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
for it in range(2):
    out = ops.BiLSTMLayerFn.apply(xp, *ps, torch.tensor(lens, dtype=torch.int32, device=DEV), int(lens.max()))
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("done")

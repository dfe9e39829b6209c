import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops, _lib
DEV='cuda:0'
B,T_,I,H = 64,350,1024,256
rng = np.random.default_rng(0)
lens = rng.integers(int(0.6*T_), T_+1, size=B); lens[0]=T_
Tp = T_+2
xp = torch.zeros((B,Tp,I), device=DEV); xp[:, :T_] = torch.randn((B,T_,I), device=DEV)
ps = [torch.empty((I+H,4*H), device=DEV).uniform_(-0.075,0.075), torch.zeros(4*H, device=DEV),
      torch.empty((I+H,4*H), device=DEV).uniform_(-0.075,0.075), torch.zeros(4*H, device=DEV)]
ops.set_gemm_mode("tf32x3")
dbg = torch.zeros(5*400, dtype=torch.int64, device=DEV)
_lib.lib().e2e_set_rec_debug(dbg.data_ptr())
for it in range(2):
    out = ops.BiLSTMLayerFn.apply(xp, *ps, torch.tensor(lens, dtype=torch.int32, device=DEV), int(lens.max()))
torch.cuda.synchronize()
d = dbg.cpu().numpy().reshape(-1,5)[:T_]
top = d[:,0]; 
print("step period (cycles): median", np.median(np.diff(top)))
w = d[1:,1]-d[1:,0]; k = d[1:,2]-d[1:,1]; pw = d[1:,3]-d[1:,2]; sy = d[1:,4]-d[1:,3]; rest = d[2:,0]-d[1:-1,4]
for name, v in [("wait",w),("kloop+reduce",k),("pointwise+stage",pw),("fence+syncthreads",sy),("issue copies+global stores->next top",rest)]:
    print("%-40s median %8.0f  p10 %8.0f p90 %8.0f" % (name, np.median(v), np.percentile(v,10), np.percentile(v,90)))

import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops, _lib
DEV='cuda:0'
np.set_printoptions(linewidth=220, precision=3, suppress=True)
M=N=128; K=32
rng = np.random.default_rng(0)
a = rng.standard_normal((M,K)).astype(np.float32); b = rng.standard_normal((N,K)).astype(np.float32)
# simple patterns: a[m,k] = m + k/100, b = identity-ish
a = (np.arange(M)[:,None] + np.arange(K)[None,:]/100.0).astype(np.float32)
b = np.zeros((N,K), np.float32); b[np.arange(32), np.arange(32)] = 1.0   # b[n,k]=delta -> C[m,n]=a[m,n] for n<32
dbg = torch.zeros(16384 + 128*128, device=DEV)
ops.ensure_workspace(DEV)
_lib.lib().e2e_set_tc_debug(dbg.data_ptr(), 1)
A = torch.tensor(a, device=DEV); B = torch.tensor(b, device=DEV)
out = ops.gemm(A, B, tb=True, mode=1)   # both K-major
torch.cuda.synchronize()
ws = ops._workspace[DEV].view(torch.float32)
print("ws A0 (big) first row:", ws[:8].cpu().numpy(), " A1 small:", ws[(M*K*4+1023)//1024*1024//4:][:8].cpu().numpy())
d = dbg.cpu().numpy()
s = d[:16384].reshape(4, 128, 32)
print("smem A_big row0:", s[0,0,:8], "row1:", s[0,1,:8], "row9", s[0,9,:12])
print("smem B_big row0:", s[2,0,:8], "row1:", s[2,1,:8])
print("smem nonzero counts:", [(np.abs(s[i])>0).sum() for i in range(4)])
tm = d[16384:].reshape(128,128)
print("tmem[0,:8]", tm[0,:8], "tmem[5,:8]", tm[5,:8], "nonzero", (np.abs(tm)>0).sum())
print("out[0,:8]", out[0,:8].cpu().numpy(), "out[5,:8]", out[5,:8].cpu().numpy())
ref = a @ b.T
print("ref[5,:8]", ref[5,:8], "max err", np.abs(out.cpu().numpy()-ref).max())

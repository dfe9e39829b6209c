import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops, _lib
DEV='cuda:0'
np.set_printoptions(linewidth=220, precision=3, suppress=True)
ops.ensure_workspace(DEV)
dbg = torch.zeros(16384 + 128*128, device=DEV)
_lib.lib().e2e_set_tc_debug(dbg.data_ptr(), 1)
rng = np.random.default_rng(0)
def run(M,N,K,ta,tb,mode=1,pattern=False):
    a = rng.standard_normal((K,M) if ta else (M,K)).astype(np.float32)
    b = rng.standard_normal((N,K) if tb else (K,N)).astype(np.float32)
    if pattern:
        # op(A)[m,k] = m + k/100 ; op(B)[k,n] = delta(k,n)
        opa = (np.arange(M)[:,None] + np.arange(K)[None,:]/100.0).astype(np.float32)
        opb = np.zeros((K,N), np.float32); opb[np.arange(min(K,N)), np.arange(min(K,N))] = 1
        a = np.ascontiguousarray(opa.T) if ta else opa
        b = np.ascontiguousarray(opb.T) if tb else opb
    dbg.zero_()
    out = ops.gemm(torch.tensor(a, device=DEV), torch.tensor(b, device=DEV), ta=ta, tb=tb, mode=mode)
    torch.cuda.synchronize()
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    o = out.cpu().numpy()
    err = np.abs(o-ref).max()/np.abs(ref).max()
    print("M,N,K=%d,%d,%d ta=%d tb=%d mode=%d relerr=%.2e  out nonzero frac %.3f" % (M,N,K,ta,tb,mode,err,(np.abs(o)>0).mean()))
    return o, ref
for ta in (0,1):
    for tb in (0,1):
        o, ref = run(128,128,32,ta,tb,pattern=True)
        if np.abs(o-ref).max() > 1e-3:
            d = dbg.cpu().numpy(); s = d[:16384].reshape(4,-1)
            print("  smem A nz %d B nz %d ; tmem nz %d" % ((np.abs(s[0])>0).sum(), (np.abs(s[2])>0).sum(), (np.abs(d[16384:])>0).sum()))
            print("  out[0,:6]", o[0,:6], "out[3,:6]", o[3,:6], " ref[3,:6]", ref[3,:6])
            print("  smemA first 40:", s[0][:40])
            print("  smemB first 40:", s[2][:40])
for (M,N,K) in [(128,128,64),(128,128,128),(128,128,264),(256,128,96),(128,256,96),(1000,520,264)]:
    for ta in (0,1):
        for tb in (0,1):
            run(M,N,K,ta,tb)
run(512,512,512,0,1,mode=2); run(512,512,512,0,0,mode=2); run(512,512,512,1,0,mode=2); run(512,512,512,1,1,mode=2)

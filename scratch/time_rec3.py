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
def report(d, NS, name, labels):
    d = d.reshape(-1, NS, 8)[:T_]
    top = d[:, 0, 0]
    print("%s NS=%d: step period (cycles) median %.0f" % (name, NS, np.median(np.diff(top[2:]))))
    for sl in range(NS):
        x = d[2:-1, sl]
        segs = ["%s %5.0f" % (labels[i], np.median(x[:, i+1]-x[:, i])) for i in range(7)]
        nxt = (d[3:, 0, 0] if sl == NS-1 else d[2:-1, sl+1, 0]) - x[:, 7]
        print("   slice %d: " % sl + " | ".join(segs) + " | loop %5.0f" % np.median(nxt))
for mode, NS in [(3, 1), (4, 2)]:
    _lib.lib().e2e_set_rec_mode(mode)
    dbg = torch.zeros(16*400, dtype=torch.int64, device=DEV)
    for it in range(2):
        out = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max()))
    torch.cuda.synchronize()
    _lib.lib().e2e_set_rec_debug(dbg.data_ptr())
    out = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max()))
    torch.cuda.synchronize()
    report(dbg.cpu().numpy().copy(), NS, "fwd", ["wait", "kloop", "combine", "pointwise", "stg+fence", "sync+mc", "gstores"])
    dbg.zero_()
    out.backward(torch.ones_like(out))
    torch.cuda.synchronize()
    report(dbg.cpu().numpy().copy(), NS, "bwd", ["wait", "sum", "pointwise+sts", "sync", "kloop", "stage+fence", "sync+send"])
    _lib.lib().e2e_set_rec_debug(0)

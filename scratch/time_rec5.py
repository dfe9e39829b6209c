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
lens_t = torch.tensor(lens, dtype=torch.int32, device=DEV)
for mode, NS in [(5, 1), (6, 2)]:
    _lib.lib().e2e_set_rec_mode(mode)
    dbg = torch.zeros(16*400, dtype=torch.int64, device=DEV)
    for it in range(2): out = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max()))
    torch.cuda.synchronize()
    _lib.lib().e2e_set_rec_debug(dbg.data_ptr())
    out = ops.BiLSTMLayerFn.apply(xp, *ps, lens_t, int(lens.max()))
    torch.cuda.synchronize()
    _lib.lib().e2e_set_rec_debug(0)
    d = dbg.cpu().numpy().reshape(-1, NS, 8)[:T_]
    x = d[5:-5]
    print("ws fwd NS=%d: M-warp step period %.0f" % (NS, np.median(np.diff(d[5:-5, 0, 0]))))
    for sl in range(NS):
        print("  slice %d  M: wait h %5.0f | kloop+z %5.0f   E: wait z %5.0f | pointwise %5.0f | publish %5.0f | stores %5.0f | M arrive -> E sees z %5.0f | E top->M next wait-done (h roundtrip) %5.0f" % (
            sl, np.median(x[:, sl, 1] - x[:, sl, 0]), np.median(x[:, sl, 2] - x[:, sl, 1]),
            np.median(x[:, sl, 4] - x[:, sl, 3]), np.median(x[:, sl, 5] - x[:, sl, 4]), np.median(x[:, sl, 6] - x[:, sl, 5]),
            np.median(x[:, sl, 7] - x[:, sl, 6]), np.median(x[:, sl, 4] - x[:, sl, 2]),
            np.median(d[6:-4, sl, 1] - x[:, sl, 6])))

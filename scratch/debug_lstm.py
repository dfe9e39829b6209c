import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops
from oracle import model as om
DEV='cuda:0'
def T(a, dtype=torch.float32): return torch.tensor(np.asarray(a), dtype=dtype, device=DEV)
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for (B,T_,I,H) in [(3,11,6,8),(17,9,40,24)]:
    rng = np.random.default_rng(B + T_ + I + H)
    lens = rng.integers(1, T_ + 1, size=B); lens[0] = T_
    x = rng.standard_normal((B, T_, I)).astype(np.float32)
    for b in range(B): x[b, lens[b]:] = 0
    ks = [rng.uniform(-0.3, 0.3, (I + H, 4 * H)).astype(np.float32) for _ in range(2)]
    bs = [rng.uniform(-0.3, 0.3, (4 * H,)).astype(np.float32) for _ in range(2)]
    ref_out, cache = om.birnn_layer_fwd(x.astype(np.float64), lens, ks[0].astype(np.float64), bs[0].astype(np.float64), ks[1].astype(np.float64), bs[1].astype(np.float64))
    Tp = T_ + 2 - (T_ % 2)
    xp = torch.zeros((B, Tp, I), device=DEV); xp[:, :T_] = T(x)
    out = ops.BiLSTMLayerFn.apply(xp, T(ks[0]), T(bs[0]), T(ks[1]), T(bs[1]), T(lens, torch.int32), int(lens.max()))
    o = out[:, :T_].cpu().numpy()
    err = np.abs(o - ref_out)
    print("lens", lens)
    print("fw err per (b,t):\n", err[:, :, :H].max(2))
    print("bw err per (b,t):\n", err[:, :, H:].max(2))
    print("fw err per unit:", err[:, :, :H].max((0,1)))

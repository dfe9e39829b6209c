"""Times the H = 512 recurrence kernels alone: ms per launch and us per timestep for 1..N batch slices (2 clusters each)."""
import sys
import torch
from e2e_asr_b200 import ops
from e2e_asr_b200._lib import call, lib

dev = torch.device("cuda:0")
H, T, nd = 512, 200, 2
Tp = T + 2
st = ops._dev_state(dev)
for B in [16, 32, 48, 56, 64, 112, 128, 256]:
    G0 = torch.randn(B * Tp, nd * 4 * H, device=dev) * 0.5
    Wh = torch.randn(nd, H, H, 4, device=dev) * 0.03
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    out = torch.zeros(B, Tp, nd * H, device=dev)
    Cst = torch.empty(B, Tp, nd, H, device=dev)
    dout = torch.randn(B, Tp, nd * H, device=dev)
    ws = ops._rec_workspace(st, B, H, nd)
    res = []
    for bwd in (False, True):
        ts = []
        for it in range(4):
            G = G0.clone()
            if bwd:
                call("e2e_lstm_rec_fwd", B, T, Tp, H, nd, Tp, 1, G, out, Cst, Wh, lens, ws, ws.numel() * 4, st["err"])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            if bwd:
                call("e2e_lstm_rec_bwd", B, T, Tp, H, nd, Tp, 1, G, Cst, Wh, dout, lens, ws, ws.numel() * 4, st["err"])
            else:
                call("e2e_lstm_rec_fwd", B, T, Tp, H, nd, Tp, 1, G, out, Cst, Wh, lens, ws, ws.numel() * 4, st["err"])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(min(ts))
    print("B=%3d clusters=%2d  fwd %.3f ms (%.2f us/step)  bwd %.3f ms (%.2f us/step)" %
          (B, nd * ((B + 15) // 16), res[0], res[0] * 1e3 / T, res[1], res[1] * 1e3 / T), flush=True)

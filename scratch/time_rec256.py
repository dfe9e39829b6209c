"""Times the H = 256 warp-specialised recurrence alone (cfg-2 layer shapes) under different rec modes."""
import torch
from e2e_asr_b200 import ops
from e2e_asr_b200._lib import call, lib

dev = torch.device("cuda:0")
H, nd, B = 256, 2, 64
st = ops._dev_state(dev)
for T in (700, 350):
    Tp = T + 2
    G0 = torch.randn(B * Tp, nd * 4 * H, device=dev) * 0.5
    Wh = torch.randn(nd, H, H, 4, device=dev) * 0.05
    lens = torch.full((B,), T, dtype=torch.int32, device=dev)
    out = torch.zeros(B, Tp, nd * H, device=dev)
    Cst = torch.empty(B, Tp, nd, H, device=dev)
    dout = torch.randn(B, Tp, nd * H, device=dev)
    ws = st["ctr"]
    for mode in (0, 0, 8):
        lib().e2e_set_rec_mode(mode)
        ts = []
        for it in range(5):
            G = G0.clone()
            call("e2e_lstm_rec_fwd", B, T, Tp, H, nd, Tp, 1, G, out, Cst, Wh, lens, ws, ws.numel() * 4, st["err"])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            call("e2e_lstm_rec_bwd", B, T, Tp, H, nd, Tp, 1, G, Cst, Wh, dout, lens, ws, ws.numel() * 4, st["err"])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("T=%d mode %d  bwd %.3f ms (%.3f us/step)  all %s" % (T, mode, min(ts), min(ts) * 1e3 / T, ["%.3f" % t for t in ts]), flush=True)
lib().e2e_set_rec_mode(0)

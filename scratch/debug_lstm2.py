import sys; sys.path.insert(0, '.')
import numpy as np, torch
from e2e_asr_b200 import ops
from e2e_asr_b200._lib import call
from oracle import model as om
DEV='cuda:0'
def T(a, dtype=torch.float32): return torch.tensor(np.asarray(a), dtype=dtype, device=DEV)
np.set_printoptions(linewidth=200, precision=4, suppress=True)
B,T_,I,H = 3,11,6,8
rng = np.random.default_rng(B + T_ + I + H)
lens = rng.integers(1, T_ + 1, size=B); lens[0] = T_
x = rng.standard_normal((B, T_, I)).astype(np.float32)
for b in range(B): x[b, lens[b]:] = 0
ks = [rng.uniform(-0.3, 0.3, (I + H, 4 * H)).astype(np.float32) for _ in range(2)]
bs = [rng.uniform(-0.3, 0.3, (4 * H,)).astype(np.float32) for _ in range(2)]
Tp = 12
xp = torch.zeros((B, Tp, I), device=DEV); xp[:, :T_] = T(x)
Wx, Wh, bp = ops._pack_lstm([T(ks[0]), T(ks[1])], [T(bs[0]), T(bs[1])], I, H, DEV)
G = ops.gemm(xp.view(B*Tp, I), Wx, bias=bp)
Gn = G.cpu().numpy().reshape(B, Tp, 2, H, 4)
# numpy reference of G
for d in range(2):
    ref = x.reshape(B*T_, I) @ ks[d][:I] + bs[d]      # [B*T, 4H] gate-blocked
    ref = ref.reshape(B, T_, 4, H).transpose(0,1,3,2)  # -> [B,T,H,4]
    print("dir", d, "G err per b:", np.abs(Gn[:, :T_, d] - ref).max((1,2,3)))
Whn = Wh.cpu().numpy()
for d in range(2):
    ref = ks[d][I:].reshape(H, 4, H).transpose(0,2,1)
    print("Wh err", np.abs(Whn[d].reshape(H,H,4) - ref).max())
out = torch.zeros((B, Tp, 2*H), device=DEV)
Cst = torch.zeros((B, Tp, 2, H), device=DEV)
st = ops._dev_state(DEV)
G0 = G.clone()
call("e2e_lstm_rec_fwd", B, int(lens.max()), Tp, H, 2, Tp, 1, G, out, Cst, Wh, T(lens, torch.int32), st["ctr"], st["ctr"].numel()*4, st["err"])
torch.cuda.synchronize()
o = out.cpu().numpy()
# expected at t=0 fw from G0
g0 = G0.cpu().numpy().reshape(B,Tp,2,H,4)
sig = lambda v: 1/(1+np.exp(-v))
for b in range(B):
    z = g0[b,0,0]
    c = sig(z[:,0])*np.tanh(z[:,1]); h = np.tanh(c)*sig(z[:,3])
    print("b",b,"expect h", h[:4], "got", o[b,0,:4], "Cst", Cst[b,0,0,:4].cpu().numpy(), "c exp", c[:4])
print("G changed rows:", (np.abs(G.cpu().numpy()-G0.cpu().numpy()).reshape(B,Tp,-1).max(2)>0).astype(int))
print("out nonzero:", (np.abs(o).max(2)>0).astype(int))

"""torch.autograd.Function wrappers around the sm_100a kernels (C ABI, _lib.py).

PyTorch is used for device-memory allocation, stream plumbing and the autograd
graph between modules only; every FLOP of the hot path runs in
libe2e_asr_b200.so.  There is no eager/CPU fallback.
"""
import contextlib
import ctypes

import numpy as np

import torch

from . import _lib
from ._lib import call

# 0 = fp32 FFMA, 1 = 3xTF32 tcgen05 (fp32-accurate), 2 = bf16 tcgen05, 3 = bf16x2 tcgen05 (hi+lo bf16, 3 products),
# 4 = f16x2 tcgen05 (hi + scaled-lo fp16, cross terms in a second accumulator: fp32-accurate at half the cost of mode 1;
#     products with a transposed A operand -- the weight gradients -- run mode 1)
_GEMM_MODE = 0
TAG_GEMM_SHAPES = False      # profiling aid: one profiler entry per GEMM shape instead of one "e2e_gemm"


def set_gemm_mode(mode):
    global _GEMM_MODE
    _GEMM_MODE = {"fp32": 0, "tf32x3": 1, "bf16": 2, "bf16x2": 3, "f16x2": 4}.get(mode, mode)


_workspace = {}


def _devkey(device):
    """Registry key of a device: always indexed ("cuda" and "cuda:0" name the same device)."""
    d = torch.device(device)
    if d.type == "cuda" and d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return str(d)


def ensure_workspace(device, nbytes=2 << 30, stream=None):
    """Registers the operand pre-pass scratch of the tensor-core GEMM modes (the default one,
    or a private one for GEMMs enqueued on a side `stream`)."""
    key = _devkey(device) if stream is None else "%s/%d" % (_devkey(device), stream.cuda_stream)
    if key not in _workspace or _workspace[key].numel() < nbytes:
        _workspace[key] = torch.empty((nbytes,), dtype=torch.uint8, device=device)
    if stream is None:
        _lib.lib().e2e_set_workspace(_workspace[key].data_ptr(), _workspace[key].numel())
    elif _lib.lib().e2e_set_stream_workspace(stream.cuda_stream, _workspace[key].data_ptr(), _workspace[key].numel()):
        raise RuntimeError("e2e_set_stream_workspace failed: %s" % _lib.lib().e2e_last_error().decode())


def get_gemm_mode():
    return _GEMM_MODE


_state = {}


def _dev_state(device):
    """Per-device scratch: recurrence step counters, barrier error flag, reduction partials."""
    key = _devkey(device)
    if key not in _state:
        _state[key] = dict(ctr=torch.zeros(1 << 20, dtype=torch.int32, device=device),
                           ctr_side=torch.zeros(1 << 20, dtype=torch.int32, device=device),
                           err=torch.zeros(1, dtype=torch.int32, device=device),
                           partials=torch.zeros(512, dtype=torch.float32, device=device),
                           one=torch.ones(1, dtype=torch.float32, device=device))
    return _state[key]


def _rec_workspace(st, B, H, nd):
    """Scratch of the recurrence kernels: the shared 4 MB buffer, or -- the H = 512 kernels reduce-scatter their partial
    d h tiles through L2, 64 KB per CTA -- a dedicated buffer grown to e2e_lstm_rec_workspace_bytes."""
    from ._lib import lib
    need = int(lib().e2e_lstm_rec_workspace_bytes(B, H, nd))
    if need <= st["ctr"].numel() * 4:
        return st["ctr"]
    big = st.get("ctr_wide")
    if big is None or big.numel() * 4 < need:
        big = st["ctr_wide"] = torch.zeros((need + 3) // 4, dtype=torch.int32, device=st["ctr"].device)
    return big


def check_device_errors(device):
    """Raises if a persistent-kernel step barrier ever timed out (host sync)."""
    st = _dev_state(device)
    if int(st["err"].item()) != 0:
        raise RuntimeError("e2e_asr_b200: a recurrence step barrier timed out on the device")


def host_array(t):
    """Host copy of a small integer tensor (lengths).  Tensors produced by
    Seq2SeqModel.get_batch carry `_host` so no device sync is needed."""
    h = getattr(t, "_host", None)
    if h is not None:
        return h
    return t.detach().cpu().numpy()


def to_i32(t, device):
    c = getattr(t, "_i32", None)
    if c is not None and c.device == torch.device(device):
        return c
    r = t.detach().to(device=device, dtype=torch.int32).contiguous()
    try:
        t._i32 = r
    except Exception:
        pass
    return r


class SplitPlanes:
    """Two-plane operand split of a tensor: `planes` is a 2-byte tensor [2, *x.shape] -- bf16x2: hi = bf16(x),
    lo = bf16(x - hi); f16x2: hi = fp16(x s), lo' = fp16((x s - hi) 2^11) with s = 1 or, for `row_inv` not None, a
    power of two per row whose inverse is row_inv[row].  Indexing takes the same view of both planes, as the caller
    indexes x (row-scaled planes: whole rows only)."""

    def __init__(self, planes, row_inv=None):
        self.planes = planes
        self.row_inv = row_inv

    @property
    def mode(self):
        return 3 if self.planes.dtype == torch.bfloat16 else 4

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        rinv = None if self.row_inv is None else self.row_inv[idx[0]]
        return SplitPlanes(self.planes[(slice(None),) + idx], rinv)

    def record_stream(self, stream):
        self.planes.record_stream(stream)
        if self.row_inv is not None:
            self.row_inv.record_stream(stream)


def split_lo(x, mode=None):
    """The operand split of a contiguous tensor for a GEMM mode (default: the current one), or None when the mode does
    not use one: tf32x3 -> x - tf32_trunc(x) (fp32, x's layout); bf16x2 / f16x2 -> SplitPlanes (f16x2: unscaled, for
    bounded operands -- weights, activations).  Views of the result, taken like the views of x, are passed to
    gemm(..., a_lo= / b_lo=) so one pass serves every product the tensor enters."""
    mode = _GEMM_MODE if mode is None else mode
    if mode not in (1, 3, 4) or x.numel() % 8 != 0 or not x.is_contiguous():
        return None
    if mode in (3, 4):
        planes = torch.empty((2,) + tuple(x.shape), dtype=torch.bfloat16 if mode == 3 else torch.float16, device=x.device)
        call("e2e_split_lo", mode, x.numel(), x, planes)
        return SplitPlanes(planes)
    lo = torch.empty_like(x)
    call("e2e_split_lo", 1, x.numel(), x, lo)
    return lo


def split_rows_f16(x):
    """f16x2 split of a contiguous 2-D gradient tensor used as the (non-transposed) A operand: every row is scaled by
    its own power of two before the split, the GEMM multiplies the output row by row_inv (e2e_split_rows_f16)."""
    if x.dim() != 2 or not x.is_contiguous() or x.shape[1] % 8 != 0:
        return None
    planes = torch.empty((2,) + tuple(x.shape), dtype=torch.float16, device=x.device)
    row_inv = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    call("e2e_split_rows_f16", x.shape[0], x.shape[1], x, planes, row_inv)
    return SplitPlanes(planes, row_inv)


def gemm(a, b, out=None, ta=False, tb=False, bias=None, z=None, accumulate=False, mode=None, a_lo=None, b_lo=None):
    """out = op(a) @ op(b) (+bias) (+z) (+out).  2-D tensors with unit inner stride;
    the leading dimension is the row stride, so column-sliced views work."""
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if ta else (a.shape[0], a.shape[1])
    Kb, N = (b.shape[1], b.shape[0]) if tb else (b.shape[0], b.shape[1])
    assert K == Kb, (a.shape, b.shape, ta, tb)
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    assert out.shape[0] == M and out.shape[1] == N and out.stride(1) == 1
    lda = a.stride(0) if a.shape[0] > 1 else max(a.shape[1], 1)
    ldb = b.stride(0) if b.shape[0] > 1 else max(b.shape[1], 1)
    ldc = out.stride(0) if out.shape[0] > 1 else max(out.shape[1], 1)
    ldz = 0
    if z is not None:
        assert z.stride(1) == 1
        ldz = z.stride(0) if z.shape[0] > 1 else z.shape[1]
    mode = _GEMM_MODE if mode is None else mode
    if mode == 4 and ta:
        mode = 1                # f16x2 scales A per output row: weight-gradient products (K over the rows) run 3xTF32
    if mode != 0 and _devkey(a.device) not in _workspace:
        ensure_workspace(a.device)
    tag = ("gemm M=%d N=%d K=%d t%d%d" % (M, N, K, ta, tb)) if TAG_GEMM_SHAPES else "e2e_gemm"
    if mode in (1, 3, 4) and (a_lo is not None or b_lo is not None):
        planes = [0, 0]
        los = [a_lo, b_lo]
        row_scale = None
        for i, (x, lo) in enumerate(((a, a_lo), (b, b_lo))):
            if lo is None:
                continue
            if (lo.mode if isinstance(lo, SplitPlanes) else 1) != mode:
                los[i] = None           # split made for another mode: let the GEMM redo its pre-pass
                continue
            if mode in (3, 4):
                if i == 0:
                    row_scale = lo.row_inv
                else:
                    assert lo.row_inv is None, "row-scaled planes serve the A operand only"
                los[i], planes[i] = lo.planes[0], lo.planes.stride(0)
            assert los[i].shape == x.shape and los[i].stride() == x.stride()
        call("e2e_gemm_lo", mode, int(ta), int(tb), M, N, K, a, los[0], lda, b, los[1], ldb, out, ldc,
             bias, z, ldz, int(accumulate), planes[0], planes[1], row_scale, work=2.0 * M * N * K, tag=tag)
        return out
    call("e2e_gemm", mode, int(ta), int(tb), M, N, K, a, lda, b, ldb, out, ldc,
         bias, z, ldz, int(accumulate), work=2.0 * M * N * K, tag=tag)
    return out


def colsum(x, out=None, accumulate=False):
    assert x.dim() == 2 and x.stride(1) == 1
    if out is None:
        out = torch.empty((x.shape[1],), dtype=torch.float32, device=x.device)
    call("e2e_colsum", x.shape[0], x.shape[1], x, x.stride(0) if x.shape[0] > 1 else x.shape[1], out,
         int(accumulate))
    return out


def flat_rows(x):
    """For a 3-D activations tensor (any of: contiguous, a time-narrowed view of a
    padded [B,Tp,D] buffer, or the transpose of one) return (flat2d, r0, r1):
    flat2d is a [rows, D] view of the underlying memory and element (i, j, :) of
    x is row i*r0 + j*r1 of it.  Lets GEMMs run over the padded buffer with no
    copy while the public tensors keep the reference's shapes."""
    assert x.dim() == 3 and x.stride(2) == 1
    D = x.shape[2]
    s0, s1 = x.stride(0), x.stride(1)
    if x.shape[0] == 1:
        s0 = x.shape[1] * s1 if s1 % D == 0 else D
    if x.shape[1] == 1:
        s1 = D
    if s0 % D or s1 % D:
        x = x.contiguous()
        s0, s1 = x.stride(0), x.stride(1)
    r0, r1 = s0 // D, s1 // D
    avail = (x.untyped_storage().nbytes() // 4 - x.storage_offset()) // D
    if r0 >= r1:
        rows = min(avail, x.shape[0] * r0)
    else:
        rows = min(avail, x.shape[1] * r1)
    need = (x.shape[0] - 1) * r0 + (x.shape[1] - 1) * r1 + 1
    assert rows >= need
    flat = torch.as_strided(x, (rows, D), (D, 1), x.storage_offset())
    return flat, r0, r1


# ---------------------------------------------------------------------------
# Encoder layer
# ---------------------------------------------------------------------------

_PREPACK = {}


def prepack_lstm(layers, device):
    """Packs the LSTM weights of `layers` (a list of (kernels, biases, I, H)) on the "pack" side stream, forked from the
    current stream: the packs depend on nothing but the parameters, so the layers above the first need not find them on
    the critical path (10 small permutation kernels, 0.36 ms per step at cfg-2).  `_pack_lstm` picks a finished pack up by
    the identity of its kernels and makes the consuming stream wait for it."""
    key = _devkey(device)
    side = _WGRAD.get(key, {}).get("pack")
    if side is None or not layers:
        return
    cur = torch.cuda.current_stream()
    fork = torch.cuda.Event()
    fork.record(cur)
    side.wait_event(fork)
    with torch.cuda.stream(side):
        for kernels, biases, I, H in layers:
            packed = _pack_lstm_now(kernels, biases, I, H, device)
            ev = torch.cuda.Event()
            ev.record(side)
            _PREPACK[(key,) + tuple(k.data_ptr() for k in kernels)] = packed + (ev,)


def _pack_lstm(kernels, biases, I, H, device):
    ent = _PREPACK.pop((_devkey(device),) + tuple(k.data_ptr() for k in kernels), None) if _PREPACK else None
    if ent is None:
        return _pack_lstm_now(kernels, biases, I, H, device)
    Wx, Wh, bp, ev = ent
    cur = torch.cuda.current_stream()
    cur.wait_event(ev)
    for t in (Wx, Wh, bp):
        t.record_stream(cur)
    return Wx, Wh, bp


def _pack_lstm_now(kernels, biases, I, H, device):
    """TF (gate-blocked) kernels of each direction -> packed Wx [I, nd*4H], Wh [nd,H,4H], bias [nd*4H]."""
    nd = len(kernels)
    Wx = torch.empty((I, nd * 4 * H), dtype=torch.float32, device=device)
    Wh = torch.empty((nd, H, 4 * H), dtype=torch.float32, device=device)
    bp = torch.empty((nd * 4 * H,), dtype=torch.float32, device=device)
    for d in range(nd):
        call("e2e_lstm_pack_weights", I, H, kernels[d], biases[d], Wx, nd * 4 * H, d * 4 * H, Wh[d], bp)
    return Wx, Wh, bp


def _unpack_lstm(dWx, dWh, dbp, I, H, nd, device):
    outs = []
    for d in range(nd):
        dk = torch.empty((I + H, 4 * H), dtype=torch.float32, device=device)
        db = torch.empty((4 * H,), dtype=torch.float32, device=device)
        call("e2e_lstm_unpack_grads", I, H, dk, db, dWx, nd * 4 * H, d * 4 * H, dWh[d], dbp, 0)
        outs += [dk, db]
    return outs


class StepSeed(int):
    """The Philox key of a step: an int (passed to the kernels by value) that may carry `.dev`, a device int64 tensor
    holding the same key.  Kernels given the device word read the key from it, so a step captured in a CUDA graph
    draws new masks at every replay: the host rewrites the word before each launch (Seq2SeqModel._stage_step_seed)."""
    dev = None

    def __new__(cls, value, dev=None):
        obj = super(StepSeed, cls).__new__(cls, int(value))
        obj.dev = dev
        return obj


def seed_dev(seed):
    return getattr(seed, "dev", None)


class DropoutFn(torch.autograd.Function):
    """DropoutWrapper(output_keep_prob) on a recurrent cell's outputs (encoder.py:50-52, decoder.py:60-63):
    y = x * mask / keep with the stateless Philox mask of e2e_dropout; backward regenerates the mask."""

    @staticmethod
    def forward(ctx, x, keep, seed, offset, first=0):
        """`first`: index of x's first element in the buffer the mask is defined over (a per-step slice of a
        [U*B, H] tensor draws that tensor's mask); a multiple of 4."""
        x = x.contiguous()
        y = torch.empty_like(x)
        call("e2e_dropout", x.numel(), x, y, float(keep), int(seed), int(offset), int(first), seed_dev(seed))
        ctx.cfg = (float(keep), int(seed), int(offset), int(first), seed_dev(seed))
        return y

    @staticmethod
    def backward(ctx, dy):
        keep, seed, offset, first, sdev = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        call("e2e_dropout", dy.numel(), dy, dx, keep, seed, offset, first, sdev)
        return dx, None, None, None, None


# Weight-gradient side stream.  The dW GEMMs of a layer (x^T dz, h^T dz, bias column sums) are off the
# backward critical path: only dX feeds the layer below.  When enabled (Seq2SeqModel does, for
# VariableStore parameters whose .grad is a view of the flat gradient buffer) they are enqueued on a
# second stream, accumulate straight into the flat buffer and overlap the next layer's latency-bound
# recurrence, which occupies 64 of the 148 SMs.  `sync_wgrad_stream` joins it before clipping.
_WGRAD = {}


def enable_wgrad_stream(device, enabled=True):
    """Two streams: "enc" (encoder layer dW GEMMs, large) and "dec" (the ~80 small launches of the decoder /
    embedding / LM-LSTM gradients), so the small launches do not queue behind the large ones."""
    key = _devkey(device)
    if not enabled:
        _WGRAD.pop(key, None)
        return None
    if key not in _WGRAD:
        _WGRAD[key] = {"enc": torch.cuda.Stream(device=device), "dec": torch.cuda.Stream(device=device),
                       "pack": torch.cuda.Stream(device=device)}
    if _GEMM_MODE != 0:
        ensure_workspace(device, nbytes=1 << 30, stream=_WGRAD[key]["enc"])
        ensure_workspace(device, nbytes=256 << 20, stream=_WGRAD[key]["dec"])
    return _WGRAD[key]


_STEP_START = {}

# Data parallelism: called as hook(grad_views, stream) when a group of parameter gradients is FINAL in the flat gradient
# buffer (all accumulation into those views has been enqueued on `stream`): the encoder layers' weight gradients on the
# "enc" stream, the decoder's on the "dec" stream.  dist.GradAllReducer starts that span's all-reduce from there, so the
# collective overlaps the rest of the backward pass.
_GRAD_READY_HOOK = None


def set_grad_ready_hook(fn):
    global _GRAD_READY_HOOK
    _GRAD_READY_HOOK = fn


def mark_step_start(device):
    """Records "the step's inputs and parameters are ready" on the current stream.  Work that depends on nothing
    else -- the decoder's LM-LSTM over the teacher-forced ids -- may start from this event on a side stream instead of
    queueing behind the encoder."""
    ev = torch.cuda.Event()
    ev.record()
    _STEP_START[_devkey(device)] = ev


_WGRAD_USED = set()      # (device key, name) of side streams that took work since the last join


def _side_stream(dev, name):
    """The "enc" / "dec" side stream of the device (None when the side streams are off).  A caller that enqueues work
    on it calls _mark_side: only streams that took work are joined by sync_wgrad_stream -- joining an idle stream
    during a CUDA-graph capture would make the captured stream wait on uncaptured work
    (cudaErrorStreamCaptureIsolation)."""
    return _WGRAD.get(_devkey(dev), {}).get(name)


def _mark_side(dev, name):
    _WGRAD_USED.add((_devkey(dev), name))


def sync_wgrad_stream(device):
    key = _devkey(device)
    ss = _WGRAD.get(key)
    if ss is not None:
        for name, s in ss.items():
            if (key, name) in _WGRAD_USED:
                torch.cuda.current_stream().wait_stream(s)
                _WGRAD_USED.discard((key, name))


class BiLSTMLayerFn(torch.autograd.Function):
    """One LSTM encoder layer over a zero-padded batch-major buffer (reference
    Encoder._layer_encoder_input, encoder.py:55-91): bidirectional (bidirectional_dynamic_rnn, fw | bw
    concatenated) or, with k_bw = b_bw = None, forward only (dynamic_rnn, bi_dir=False).

    x: [B, Tp, I] contiguous with Tp >= max(len)+1; returns [B, Tp, nd*H], zero for t >= len.
    The pyramid (encoder.py:94-119) is then a free reshape.
    """

    @staticmethod
    def forward(ctx, x, k_fw, b_fw, k_bw, b_bw, lens_i32, T):
        B, Tp, I = x.shape
        H = k_fw.shape[1] // 4
        nd = 1 if k_bw is None else 2
        dev = x.device
        assert x.is_contiguous() and Tp >= T + 1
        st = _dev_state(dev)
        Wx, Wh, bp = _pack_lstm([k_fw, k_bw][:nd], [b_fw, b_bw][:nd], I, H, dev)
        x2 = x.view(B * Tp, I)
        # 3xTF32 "small" halves, computed once per tensor and shared by every product it enters (fwd + bwd)
        x_lo, Wx_lo = split_lo(x2), split_lo(Wx)
        G = gemm(x2, Wx, bias=bp, a_lo=x_lo, b_lo=Wx_lo)            # [B*Tp, nd*4H]
        out = torch.zeros((B, Tp, nd * H), dtype=torch.float32, device=dev)
        Cst = torch.empty((B, Tp, nd, H), dtype=torch.float32, device=dev)
        ws = _rec_workspace(st, B, H, nd)
        call("e2e_lstm_rec_fwd", B, T, Tp, H, nd, Tp, 1, G, out, Cst, Wh, lens_i32, ws,
             ws.numel() * 4, st["err"], work=float(T), tag="enc_rec_fwd")
        ctx.save_for_backward(x, Wx, Wh, G, Cst, out, lens_i32)
        ctx.los = (x_lo, Wx_lo)
        ctx.dims = (B, Tp, I, H, T, nd)
        # flat-gradient-buffer views of the parameters (None for plain tensors)
        ctx.grad_dst = tuple(getattr(t, "grad", None) for t in (k_fw, b_fw, k_bw, b_bw)[:2 * nd])
        return out

    @staticmethod
    def backward(ctx, dout):
        x, Wx, Wh, G, Cst, out, lens_i32 = ctx.saved_tensors
        B, Tp, I, H, T, nd = ctx.dims
        dev = x.device
        st = _dev_state(dev)
        dout = dout.contiguous()
        ws = _rec_workspace(st, B, H, nd)
        call("e2e_lstm_rec_bwd", B, T, Tp, H, nd, Tp, 1, G, Cst, Wh, dout, lens_i32, ws,
             ws.numel() * 4, st["err"], work=float(T), tag="enc_rec_bwd")   # G now holds d(pre-activations)
        N = B * Tp
        x2, o2 = x.view(N, I), out.view(N, nd * H)
        x_lo, Wx_lo = ctx.los
        f16 = _GEMM_MODE == 4
        if f16:
            # f16x2: the critical-path product dX = dz . Wx^T takes dz as fp16 planes scaled per row (gradients span
            # many orders of magnitude across utterances) and Wx's planes of the forward pass; the weight-gradient
            # products (K runs over the rows) run 3xTF32 on the side stream with their own splits
            G_lo = split_rows_f16(G)
        else:
            G_lo = split_lo(G)        # one split of dz serves the dX, dW_x and the dW_h products
        dX = gemm(G, Wx, tb=True, a_lo=G_lo, b_lo=Wx_lo).view(B, Tp, I) if ctx.needs_input_grad[0] else None

        def weight_grads():
            xl, gl = (split_lo(x2, 1), split_lo(G, 1)) if f16 else (x_lo, G_lo)
            dWx = gemm(x2, G, ta=True, a_lo=xl, b_lo=gl)                # [I, nd*4H]
            dWh = torch.empty((nd, H, 4 * H), dtype=torch.float32, device=dev)
            # h_{t-1}^T dz_t: fw pairs out[t-1] with dz[t], bw pairs out[t+1] with dz[t]; the flat
            # one-row shift never crosses an utterance because out[b, Tp-1] == 0 and dz[b, Tp-1] == 0.
            gemm(o2[:N - 1, 0:H], G[1:, 0:4 * H], ta=True, out=dWh[0],
                 b_lo=None if gl is None else gl[1:, 0:4 * H])
            if nd == 2:
                gemm(o2[1:, H:2 * H], G[:N - 1, 4 * H:8 * H], ta=True, out=dWh[1],
                     b_lo=None if gl is None else gl[:N - 1, 4 * H:8 * H])
            dbp = colsum(G)
            return dWx, dWh, dbp

        side = _side_stream(dev, "enc")
        dst = ctx.grad_dst
        pad = (None,) * (2 * (2 - nd))
        if side is not None and all(d is not None for d in dst) and all(ctx.needs_input_grad[1:1 + 2 * nd]):
            main = torch.cuda.current_stream()
            _mark_side(dev, "enc")
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dWx, dWh, dbp = weight_grads()
                for d in range(nd):     # += into the flat gradient buffer (what AccumulateGrad would do)
                    call("e2e_lstm_unpack_grads", I, H, dst[2 * d], dst[2 * d + 1], dWx, nd * 4 * H, d * 4 * H,
                         dWh[d], dbp, 1)
            for t_ in (x, G, out, x_lo, G_lo):
                if t_ is not None:
                    t_.record_stream(side)
            if _GRAD_READY_HOOK is not None:
                _GRAD_READY_HOOK(dst, side)
            return dX, None, None, None, None, None, None
        dWx, dWh, dbp = weight_grads()
        return (dX,) + tuple(_unpack_lstm(dWx, dWh, dbp, I, H, nd, dev)) + pad + (None, None)


class BiGRULayerFn(torch.autograd.Function):
    """One GRU encoder layer (reference encoder.py:48 `GRUCell` when use_lstm=False, under
    (bidirectional_)dynamic_rnn, encoder.py:77-89) over a zero-padded batch-major buffer.

    x: [B, Tp, I]; per direction the TF variables gates/{kernel [(I+H), 2H], bias [2H]} and
    candidate/{kernel [(I+H), H], bias [H]}; returns [B, Tp, nd*H] (fw | bw), zero for t >= len.
    The x halves of both kernels are batched GEMMs over all frames, the recurrence is e2e_gru_rec_fwd/bwd (one CTA
    per 4 utterances, no inter-CTA synchronisation); the parameter gradients are GEMMs over the stored
    pre-activation gradients, returned through autograd."""

    @staticmethod
    def forward(ctx, x, lens_i32, T, *params):
        B, Tp, I = x.shape
        nd = len(params) // 4
        H = params[2].shape[1]
        dev = x.device
        assert x.is_contiguous() and Tp >= T + 1 and len(params) in (4, 8)
        gk, gb, ck, cb = (params[0::4], params[1::4], params[2::4], params[3::4])
        Wg_x = torch.cat([k[:I] for k in gk], dim=1).contiguous()          # [I, nd*2H]
        Wc_x = torch.cat([k[:I] for k in ck], dim=1).contiguous()          # [I, nd*H]
        Wg_h = torch.stack([k[I:] for k in gk]).contiguous()               # [nd, H, 2H]
        Wc_h = torch.stack([k[I:] for k in ck]).contiguous()               # [nd, H, H]
        x2 = x.view(B * Tp, I)
        Gg = gemm(x2, Wg_x, bias=torch.cat(list(gb)))
        Gc = gemm(x2, Wc_x, bias=torch.cat(list(cb)))
        out = torch.zeros((B, Tp, nd * H), dtype=torch.float32, device=dev)
        RH = torch.zeros((B * Tp, nd * H), dtype=torch.float32, device=dev)
        call("e2e_gru_rec_fwd", B, T, Tp, H, nd, Gg, Gc, out, RH, Wg_h, Wc_h, lens_i32, work=float(T),
             tag="enc_gru_fwd")
        ctx.save_for_backward(x, Wg_x, Wc_x, Wg_h, Wc_h, Gg, Gc, out, RH, lens_i32)
        ctx.dims = (B, Tp, I, H, T, nd)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, Wg_x, Wc_x, Wg_h, Wc_h, Gg, Gc, out, RH, lens_i32 = ctx.saved_tensors
        B, Tp, I, H, T, nd = ctx.dims
        dout = dout.contiguous()
        call("e2e_gru_rec_bwd", B, T, Tp, H, nd, Gg, Gc, out, dout, Wg_h.transpose(1, 2).contiguous(),
             Wc_h.transpose(1, 2).contiguous(), lens_i32, work=float(T), tag="enc_gru_bwd")
        # Gg / Gc now hold d(gate pre-activations) / d(candidate pre-activation), zero past each length
        N = B * Tp
        x2, o2 = x.view(N, I), out.view(N, nd * H)
        dX = None
        if ctx.needs_input_grad[0]:
            dX = gemm(Gg, Wg_x, tb=True)
            gemm(Gc, Wc_x, tb=True, out=dX, accumulate=True)
            dX = dX.view(B, Tp, I)
        dWg_x = gemm(x2, Gg, ta=True)                                      # [I, nd*2H]
        dWc_x = gemm(x2, Gc, ta=True)                                      # [I, nd*H]
        dbg, dbc = colsum(Gg), colsum(Gc)
        grads = []
        for d in range(nd):
            cg, cc = slice(d * 2 * H, (d + 1) * 2 * H), slice(d * H, (d + 1) * H)
            # h_{t-1}^T d(gates)_t: fw pairs out[t-1] with row t, bw pairs out[t+1] with row t (one-row shift of the
            # flat buffer; it never crosses an utterance because the last padded frame of every utterance is zero)
            if d == 0:
                dWg_h = gemm(o2[:N - 1, cc], Gg[1:, cg], ta=True)
            else:
                dWg_h = gemm(o2[1:, cc], Gg[:N - 1, cg], ta=True)
            dWc_h = gemm(RH[:, cc], Gc[:, cc], ta=True)                    # (r * h_{t-1})^T d(candidate)
            grads += [torch.cat([dWg_x[:, cg], dWg_h], dim=0), dbg[cg].clone(),
                      torch.cat([dWc_x[:, cc], dWc_h], dim=0), dbc[cc].clone()]
        return (dX, None, None) + tuple(grads)


# ---------------------------------------------------------------------------
# General decoder cells (decoder.py:49-82): MultiRNNCell stacks (num_layers_dec > 1) and GRU cells (use_lstm=False).
# Autograd composes single steps; every step's arithmetic is an e2e_gemm or a pointwise kernel of cell_point.cu.
# A functional path: the benchmarked single-layer LSTM decoder runs the persistent kernels (AttnDecoderFnV2).
# ---------------------------------------------------------------------------

class LinearFn(torch.autograd.Function):
    """tf `_linear(x, n_out, True)`: x . W + b."""

    @staticmethod
    def forward(ctx, x, W, b):
        x = x.contiguous()
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return gemm(x, W, bias=b)

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        dy = dy.contiguous()
        return gemm(dy, W, tb=True), gemm(x, dy, ta=True), colsum(dy) if ctx.has_bias else None


class LSTMPointFn(torch.autograd.Function):
    """BasicLSTMCell after the matmul: (z [n,4H], c) -> (c', h')."""

    @staticmethod
    def forward(ctx, z, c_prev):
        z, c_prev = z.contiguous(), c_prev.contiguous()
        n, H = c_prev.shape
        c_new, h_new = torch.empty_like(c_prev), torch.empty_like(c_prev)
        call("e2e_lstm_point_fwd", n, H, z, c_prev, c_new, h_new)
        ctx.save_for_backward(z, c_prev, c_new)
        return c_new, h_new

    @staticmethod
    def backward(ctx, dc_new, dh_new):
        z, c_prev, c_new = ctx.saved_tensors
        n, H = c_prev.shape
        dz, dc_prev = torch.empty_like(z), torch.empty_like(c_prev)
        call("e2e_lstm_point_bwd", n, H, z, c_prev, c_new, None if dc_new is None else dc_new.contiguous(),
             None if dh_new is None else dh_new.contiguous(), dz, dc_prev)
        return dz, dc_prev


class GRUGateFn(torch.autograd.Function):
    """GRUCell gates: (zg [n,2H] = (r|u), h) -> (sig(r) * h, sig(u))."""

    @staticmethod
    def forward(ctx, zg, h_prev):
        zg, h_prev = zg.contiguous(), h_prev.contiguous()
        n, H = h_prev.shape
        rh, u = torch.empty_like(h_prev), torch.empty_like(h_prev)
        call("e2e_gru_gate_fwd", n, H, zg, h_prev, rh, u)
        ctx.save_for_backward(zg, h_prev)
        return rh, u

    @staticmethod
    def backward(ctx, drh, du):
        zg, h_prev = ctx.saved_tensors
        n, H = h_prev.shape
        dzg, dh_prev = torch.empty_like(zg), torch.empty_like(h_prev)
        call("e2e_gru_gate_bwd", n, H, zg, h_prev, None if drh is None else drh.contiguous(),
             None if du is None else du.contiguous(), dzg, dh_prev)
        return dzg, dh_prev


class GRUOutFn(torch.autograd.Function):
    """GRUCell output: h' = u h + (1 - u) tanh(zc)."""

    @staticmethod
    def forward(ctx, zc, u, h_prev):
        zc, u, h_prev = zc.contiguous(), u.contiguous(), h_prev.contiguous()
        n, H = h_prev.shape
        h_new = torch.empty_like(h_prev)
        call("e2e_gru_out_fwd", n, H, zc, u, h_prev, h_new)
        ctx.save_for_backward(zc, u, h_prev)
        return h_new

    @staticmethod
    def backward(ctx, dh_new):
        zc, u, h_prev = ctx.saved_tensors
        n, H = h_prev.shape
        dzc, du, dh_prev = torch.empty_like(zc), torch.empty_like(u), torch.empty_like(h_prev)
        call("e2e_gru_out_bwd", n, H, zc, u, h_prev, dh_new.contiguous(), dzc, du, dh_prev)
        return dzc, du, dh_prev


class EmbedFn(torch.autograd.Function):
    """embedding_lookup (decoder.py:101) with its IndexedSlices gradient (scatter-add; the per-occurrence rows are kept
    in `stash["emb_values"]` for tf.global_norm's view of them)."""

    @staticmethod
    def forward(ctx, emb, ids, stash):
        ids = ids.contiguous()
        n, E = ids.numel(), emb.shape[1]
        out = torch.empty((n, E), dtype=torch.float32, device=emb.device)
        call("e2e_embed_gather", n, E, emb, ids, out)
        ctx.save_for_backward(ids)
        ctx.shape, ctx.stash = emb.shape, stash
        return out

    @staticmethod
    def backward(ctx, dout):
        (ids,) = ctx.saved_tensors
        dout = dout.contiguous()
        demb = torch.zeros(ctx.shape, dtype=torch.float32, device=dout.device)
        call("e2e_embed_scatter_add", ids.numel(), ctx.shape[1], demb, ids, dout, dout.shape[1])
        if ctx.stash is not None:
            ctx.stash["emb_values"] = dout
        return demb, None, None


class AttnStepFn(torch.autograd.Function):
    """attention() of attn_decoder.py:77-93 for one step: (y [B,A], HF [B*Tp,A], enc [B*Tp,D], v) -> ctx [B,D]."""

    @staticmethod
    def forward(ctx, y, HF, enc_flat, v, enc_len_i32, dims):
        B, Tn, Tp, A, D = dims
        y = y.contiguous()
        alpha = torch.empty((B, Tn), dtype=torch.float32, device=y.device)
        out = torch.empty((B, D), dtype=torch.float32, device=y.device)
        call("e2e_attn_fwd", B, Tn, Tp, A, D, HF, enc_flat, enc_len_i32, y, v, alpha, out, D)
        ctx.save_for_backward(y, HF, enc_flat, v, enc_len_i32, alpha)
        ctx.dims = dims
        return out

    @staticmethod
    def backward(ctx, dctx):
        y, HF, enc_flat, v, enc_len_i32, alpha = ctx.saved_tensors
        B, Tn, Tp, A, D = ctx.dims
        dctx = dctx.contiguous()
        dHF, denc = torch.zeros_like(HF), torch.zeros_like(enc_flat)
        dy = torch.empty_like(y)
        dv_part = torch.zeros((B, A), dtype=torch.float32, device=y.device)
        call("e2e_attn_bwd", B, Tn, Tp, A, D, HF, enc_flat, enc_len_i32, y, v, alpha, dctx, D, dHF, denc, dy, dv_part)
        return dy, dHF, denc, colsum(dv_part), None, None


def _cell_step(use_lstm, x, state, ws):
    """One TF cell step on rows: LSTM state = (c, h), GRU state = (h,).  Returns (output h', new state)."""
    if use_lstm:
        c, h = state
        c2, h2 = LSTMPointFn.apply(LinearFn.apply(torch.cat([x, h], dim=1), ws[0], ws[1]), c)
        return h2, (c2, h2)
    (h,) = state
    rh, u = GRUGateFn.apply(LinearFn.apply(torch.cat([x, h], dim=1), ws[0], ws[1]), h)
    h2 = GRUOutFn.apply(LinearFn.apply(torch.cat([x, rh], dim=1), ws[2], ws[3]), u, h)
    return h2, (h2,)


def attn_decoder_stepwise(enc, v, lm_cells, dec_cells, use_lstm, ids, lens_i32, enc_len_i32, U, stash=None,
                          drop=None, feedback=None):
    """AttnDecoder.__call__ (attn_decoder.py:37-172) under teacher forcing for ANY decoder.py cell configuration:
    `lm_cells` / `dec_cells` are lists (one entry per stacked layer) of the cell's variables -- (kernel, bias) for
    BasicLSTMCell, (gates kernel, gates bias, candidate kernel, candidate bias) for GRUCell.  The attention query and
    the projection input are get_state(state): the last layer's c (LSTM) or state (GRU), decoder.py:74-82; raw_rnn
    copies the decoder state through for finished rows and zeroes their emit; the lm state is never frozen.
    drop = (keep, seed, task index) in training with out_prob_dec < 1: DropoutWrapper(output_keep_prob) on every
    single cell (decoder.py:60-63) -- each layer's OUTPUT is dropped on its way to the next layer / to
    InputProjection, the states are not; the top decoder layer's output is never read.  Philox streams: 100 + task
    for the top lm layer (as for the single cell), 500 + 16 task + l for lower lm layers, 400 + 16 task + l for the
    decoder layers; the mask of step t is the slice [t*B, (t+1)*B) of a [U*B, H] mask.
    feedback: None = teacher forcing (step t reads ids[t]); "greedy" = eval mode, step t+1 embeds argmax of step t's
    emit (decoder.py:139-153); a dict(use_sample, seed, offset, ids) = the scheduled-sampling rule of
    inference.sample_decode_ids: after step t the input of step t+1 is written to feedback["ids"][t+1] (run without a
    tape; the training step then takes the teacher-forced path on the realised ids).
    Returns logits [(U*B), V]."""
    dev = enc.device
    B, Tn, D = enc.shape
    A = v["q_k"].shape[1]
    enc_flat, Tp = enc.contiguous().view(B * Tn, D), Tn          # plain differentiable copy: autograd composes this path
    f32 = dict(dtype=torch.float32, device=dev)
    Hl = lm_cells[0][-1].shape[0] // (4 if use_lstm else 1)
    Hd = dec_cells[0][-1].shape[0] // (4 if use_lstm else 1)
    HF = LinearFn.apply(enc_flat, v["attn_w"].view(D, A), None)
    V = v["out_k"].shape[1]
    u_all = EmbedFn.apply(v["emb"], ids[:U], stash).view(U, B, -1) if feedback is None else None
    tok = ids[0].contiguous()
    zero = (lambda H: (torch.zeros((B, H), **f32), torch.zeros((B, H), **f32))) if use_lstm else \
        (lambda H: (torch.zeros((B, H), **f32),))
    lm_state = [zero(Hl) for _ in lm_cells]
    dec_state = [zero(Hd) for _ in dec_cells]
    ctx_vec = torch.zeros((B, D), **f32)
    steps_t = torch.arange(U, device=dev, dtype=torch.int32)
    live_all = (steps_t[:, None] < lens_i32[None, :]).unsqueeze(2)           # [U, B, 1]
    dims = (B, Tn, Tp, A, D)
    outs = []
    for t in range(U):
        x = u_all[t] if feedback is None else EmbedFn.apply(v["emb"], tok, None)
        for l, ws in enumerate(lm_cells):                                     # lm_cell stack (attn_decoder.py:148)
            x, lm_state[l] = _cell_step(use_lstm, x, lm_state[l], ws)
            if drop is not None:
                stream = 100 + drop[2] if l == len(lm_cells) - 1 else 500 + 16 * drop[2] + l
                x = DropoutFn.apply(x, drop[0], drop[1], stream, t * B * Hl)
        m = LinearFn.apply(x, v["sp_k"], v["sp_b"]) if v["sp_k"] is not None else x
        x = LinearFn.apply(torch.cat([m, ctx_vec], dim=1), v["in_k"], v["in_b"])       # InputProjection (:157-158)
        new_state = []
        for l, ws in enumerate(dec_cells):                                    # decoder cell stack (raw_rnn body)
            x, st_new = _cell_step(use_lstm, x, dec_state[l], ws)
            new_state.append(st_new)
            if drop is not None and l < len(dec_cells) - 1:
                x = DropoutFn.apply(x, drop[0], drop[1], 400 + 16 * drop[2] + l, t * B * Hd)
        q = new_state[-1][0]
        y = LinearFn.apply(q, v["q_k"], v["q_b"])
        ctx_vec = AttnStepFn.apply(y, HF, enc_flat, v["attn_v"], enc_len_i32, dims)
        proj = LinearFn.apply(torch.cat([q, ctx_vec], dim=1), v["ap_k"], v["ap_b"])
        lg = LinearFn.apply(proj, v["out_k"], v["out_b"])
        live = live_all[t]
        outs.append(torch.where(live, lg, torch.zeros_like(lg)))
        dec_state = [tuple(torch.where(live, a_, b_) for a_, b_ in zip(new_state[l], dec_state[l]))
                     for l in range(len(dec_cells))]
        if feedback == "greedy":                       # argmax of the EMIT (zero for finished rows), decoder.py:148-151
            tok = torch.empty((B,), dtype=torch.int64, device=dev)
            call("e2e_argmax_rows", B, V, outs[-1].contiguous(), V, tok)
        elif feedback is not None and t + 1 < feedback["ids"].shape[0]:
            tok = feedback["ids"][t + 1]
            if feedback["use_sample"][t + 1]:          # multinomial over the previous step's (unmasked) logits
                call("e2e_sample_rows", B, V, lg.contiguous(), V, int(feedback["seed"]), int(feedback["offset"]),
                     (t + 1) * B, tok)
    return torch.cat(outs, dim=0)


# ---------------------------------------------------------------------------
# Attention decoder (teacher forced)
# ---------------------------------------------------------------------------

class AttnDecoderFn(torch.autograd.Function):
    """AttnDecoder.__call__ under teacher forcing (attn_decoder.py:37-172; step
    order SURVEY.md A.4).  Everything that does not depend on the decoder state is
    batched over all U steps (embedding, LM-LSTM, lm half of InputProjection,
    hidden_features, AttnProjection, OutputProjection); only
    xin -> dec-LSTM -> attention runs step by step (e2e_decoder_loop_*)."""

    @staticmethod
    def forward(ctx, enc, emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, q_b, ap_k, ap_b, out_k, out_b,
                in_k, in_b, sp_k, sp_b, ids, lens_i32, enc_len_i32, U, stash, lm_drop=None):
        dev = enc.device
        st = _dev_state(dev)
        B, Tn, D = enc.shape
        V, E = emb.shape
        Hl, Hd, A = lm_k.shape[1] // 4, dec_k.shape[1] // 4, q_k.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        enc_flat, r0, r1 = flat_rows(enc)
        assert r1 == 1, "encoder states must be batch-major"
        Tp = r0
        ids = ids[:U].contiguous()
        # --- state-independent, batched over all steps
        u = torch.empty((U * B, E), **f32)
        call("e2e_embed_gather", U * B, E, emb, ids, u)
        Wx_lm, Wh_lm, bp_lm = _pack_lstm([lm_k], [lm_b], E, Hl, dev)
        G_lm = gemm(u, Wx_lm, bias=bp_lm)                            # [U*B, 4Hl]
        hl = torch.zeros((U * B, Hl), **f32)
        C_lm = torch.empty((U * B, Hl), **f32)
        call("e2e_lstm_rec_fwd", B, U, U, Hl, 1, 1, B, G_lm, hl, C_lm, Wh_lm, lens_i32, st["ctr"],
             st["ctr"].numel() * 4, st["err"], work=float(U), tag="lm_rec_fwd")
        # DropoutWrapper on lm_cell: its OUTPUT is dropped, the recurrent state is not (decoder.py:60-63)
        hl_out = hl
        if lm_drop is not None:
            hl_out = torch.empty_like(hl)
            call("e2e_dropout", hl.numel(), hl, hl_out, float(lm_drop[0]), int(lm_drop[1]), int(lm_drop[2]), 0,
                 seed_dev(lm_drop[1]))
        m = gemm(hl_out, sp_k, bias=sp_b) if sp_k is not None else hl_out   # SimpleProjection (:149-151)
        pre = gemm(m, in_k[:Hd], bias=in_b)                          # lm half of InputProjection (:157-158)
        HF = gemm(enc_flat, attn_w.view(D, A))                       # hidden_features (:70-73), padded rows too
        # --- sequential loop
        a = _lib.DecLoopFwdArgs()
        a.B, a.U, a.E, a.Hd, a.A, a.D, a.Tn, a.Tp, a.gemm_mode = B, U, E, Hd, A, D, Tn, Tp, _GEMM_MODE
        bufs = dict(xh=torch.zeros((U * B, E + Hd), **f32), cprev=torch.zeros((U * B, Hd), **f32),
                    acts=torch.empty((U * B, 4 * Hd), **f32), cat=torch.empty((U * B, Hd + D), **f32),
                    y=torch.empty((U * B, A), **f32), alpha=torch.empty((U * B, Tn), **f32),
                    gates_tmp=torch.empty((B, 4 * Hd), **f32))
        ptrs = dict(in_k=in_k, dec_k=dec_k, dec_b=dec_b, q_k=q_k, q_b=q_b, attn_v=attn_v, pre=pre, HF=HF,
                    enc=enc_flat, enc_len=enc_len_i32, lens=lens_i32, **bufs)
        for k, v in ptrs.items():
            setattr(a, k, v.data_ptr())
        call("e2e_decoder_loop_fwd", a, work=float(U))
        proj = gemm(bufs["cat"], ap_k, bias=ap_b)                     # AttnProjection (:116-118)
        logits = gemm(proj, out_k, bias=out_b)                        # OutputProjection (:124-125)
        call("e2e_mask_rows", U, B, V, logits, lens_i32)              # raw_rnn zeroes finished rows
        ctx.save_for_backward(enc, emb, attn_w, attn_v, lm_k, dec_k, dec_b, q_k, q_b, ap_k, out_k, in_k, sp_k,
                              ids, lens_i32, enc_len_i32, u, Wx_lm, Wh_lm, G_lm, hl, C_lm, m, pre, HF, proj,
                              bufs["xh"], bufs["cprev"], bufs["acts"], bufs["cat"], bufs["y"], bufs["alpha"],
                              bufs["gates_tmp"])
        ctx.dims = (B, Tn, D, V, E, Hl, Hd, A, U, Tp)
        ctx.stash = stash
        ctx.lm_drop = lm_drop
        ctx.hl_out = hl_out if lm_drop is not None else None
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        (enc, emb, attn_w, attn_v, lm_k, dec_k, dec_b, q_k, q_b, ap_k, out_k, in_k, sp_k, ids, lens_i32,
         enc_len_i32, u, Wx_lm, Wh_lm, G_lm, hl, C_lm, m, pre, HF, proj, xh, cprev, acts, cat, y, alpha,
         gates_tmp) = ctx.saved_tensors
        B, Tn, D, V, E, Hl, Hd, A, U, Tp = ctx.dims
        dev = enc.device
        st = _dev_state(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        dlogits = dlogits.contiguous()
        enc_flat, _, _ = flat_rows(enc)
        # batched projections
        dout_k = gemm(proj, dlogits, ta=True)
        dout_b = colsum(dlogits)
        dproj = gemm(dlogits, out_k, tb=True)
        dap_k = gemm(cat, dproj, ta=True)
        dap_b = colsum(dproj)
        dcat = gemm(dproj, ap_k, tb=True)                             # [U*B, Hd+D]
        # sequential loop
        g = _lib.DecLoopBwdArgs()
        a = g.f
        a.B, a.U, a.E, a.Hd, a.A, a.D, a.Tn, a.Tp, a.gemm_mode = B, U, E, Hd, A, D, Tn, Tp, _GEMM_MODE
        for k, v in dict(in_k=in_k, dec_k=dec_k, dec_b=dec_b, q_k=q_k, q_b=q_b, attn_v=attn_v, pre=pre, HF=HF,
                         enc=enc_flat, enc_len=enc_len_i32, lens=lens_i32, xh=xh, cprev=cprev, acts=acts,
                         cat=cat, y=y, alpha=alpha, gates_tmp=gates_tmp).items():
            setattr(a, k, v.data_ptr())
        nrows = enc_flat.shape[0]
        out = dict(dcat=dcat, dgates=torch.empty((U * B, 4 * Hd), **f32), dxh=torch.empty((U * B, E + Hd), **f32),
                   dy=torch.empty((U * B, A), **f32), dv_part=torch.empty((B * Tn, A), **f32),
                   dHF=torch.zeros((nrows, A), **f32), denc=torch.zeros((nrows, D), **f32),
                   dc_carry=torch.zeros((B, Hd), **f32), ds=torch.empty((U * B, Tn), **f32))
        for k, v in out.items():
            setattr(g, k, v.data_ptr())
        call("e2e_decoder_loop_bwd", g, work=float(U))
        dgates, dxh, dy, dHF, denc = out["dgates"], out["dxh"], out["dy"], out["dHF"], out["denc"]
        dq_k = gemm(cat[:, :Hd], dy, ta=True)
        dq_b = colsum(dy)
        dattn_v = colsum(out["dv_part"])
        ddec_k = gemm(xh, dgates, ta=True)
        ddec_b = colsum(dgates)
        dxin = dxh[:, :E]
        din_k = torch.zeros((Hd + D, E), **f32)
        gemm(m, dxin, ta=True, out=din_k[:Hd])
        if U > 1:   # ctx_{t-1} pairs with dxin_t: time-major rows shift by B
            gemm(cat[:(U - 1) * B, Hd:], dxin[B:], ta=True, out=din_k[Hd:])
        din_b = colsum(dxin)
        dm = gemm(dxin, in_k[:Hd], tb=True)                           # [U*B, Hd]
        dsp_k = dsp_b = None
        if sp_k is not None:
            dsp_k = gemm(hl if ctx.hl_out is None else ctx.hl_out, dm, ta=True)
            dsp_b = colsum(dm)
            dm = gemm(dm, sp_k, tb=True)
        if ctx.lm_drop is not None:
            dmd = torch.empty_like(dm)
            call("e2e_dropout", dm.numel(), dm.contiguous(), dmd, float(ctx.lm_drop[0]), int(ctx.lm_drop[1]),
                 int(ctx.lm_drop[2]), 0, seed_dev(ctx.lm_drop[1]))
            dm = dmd
        # LM-LSTM backward (time-major rows: b stride 1, t stride B)
        call("e2e_lstm_rec_bwd", B, U, U, Hl, 1, 1, B, G_lm, C_lm, Wh_lm, dm, lens_i32, st["ctr"],
             st["ctr"].numel() * 4, st["err"], work=float(U), tag="lm_rec_bwd")
        dWx_lm = gemm(u, G_lm, ta=True)
        dWh_lm = torch.zeros((1, Hl, 4 * Hl), **f32)
        if U > 1:
            gemm(hl[:(U - 1) * B], G_lm[B:], ta=True, out=dWh_lm[0])
        dbp_lm = colsum(G_lm)
        dlm_k, dlm_b = _unpack_lstm(dWx_lm, dWh_lm, dbp_lm, E, Hl, 1, dev)
        du = gemm(G_lm, Wx_lm, tb=True)                               # [U*B, E] = IndexedSlices values
        demb = torch.zeros((V, E), **f32)
        call("e2e_embed_scatter_add", U * B, E, demb, ids, du, E)
        if ctx.stash is not None:
            # tf.global_norm takes the embedding gradient's IndexedSlices values
            ctx.stash["emb_values"] = du
        # hidden_features = enc (*) AttnW
        dattn_w = gemm(enc_flat, dHF, ta=True).view(attn_w.shape)
        gemm(dHF, attn_w.view(D, A), tb=True, out=denc, accumulate=True)
        denc_view = torch.as_strided(denc, (B, Tn, D), (Tp * D, D, 1)) if ctx.needs_input_grad[0] else None
        return (denc_view, demb, dattn_w, dattn_v, dlm_k, dlm_b, ddec_k, ddec_b, dq_k, dq_b, dap_k, dap_b, dout_k,
                dout_b, din_k, din_b, dsp_k, dsp_b, None, None, None, None, None, None)


_DECODER_IMPL = "persist"      # "persist": one cooperative launch per direction; "loop": per-step kernels


def set_decoder_impl(name):
    global _DECODER_IMPL
    assert name in ("persist", "loop")
    _DECODER_IMPL = name


def _persist_fits(enc, dec_k, q_k, U):
    a = _lib.DecPersistArgs()
    a.B, a.Tn, a.D = enc.shape
    a.U, a.Hd, a.A, a.Tp = int(U), dec_k.shape[1] // 4, q_k.shape[1], flat_rows(enc)[1]
    return bool(_lib.lib().e2e_decoder_persist_fits(ctypes.addressof(a)))


def attn_decoder_apply(*args, lm_drop=None, samp=None):
    # args: enc, emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, ...
    if _DECODER_IMPL == "persist" and _persist_fits(args[0], args[6], args[8], args[21]):
        return AttnDecoderFnV2.apply(*(args + (lm_drop, samp)))
    assert samp is None, "in-pass scheduled sampling is served by the persistent decoder kernels"
    return AttnDecoderFn.apply(*(args + (lm_drop,)))


class AttnDecoderFnV2(torch.autograd.Function):
    """Same contract as AttnDecoderFn, built on the persistent decoder kernels
    (csrc/decoder_persist.cu).  InputProjection's ctx half is folded into the
    decoder-LSTM kernel:  gates_t = pre_g[t] + [ctx_{t-1} | h_{t-1}] . W_ch with
    W_ch = [in_k[Hd:] . Wx ; Wh] and pre_g = (m . in_k[:Hd] + in_b) . Wx + b, so the
    sequential loop is gates -> attention only; the chain rule through the product
    W_cx = in_k[Hd:] . Wx is applied after the loop."""

    @staticmethod
    def forward(ctx, enc, emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, q_b, ap_k, ap_b, out_k, out_b,
                in_k, in_b, sp_k, sp_b, ids, lens_i32, enc_len_i32, U, stash, lm_drop=None, samp=None):
        dev = enc.device
        st = _dev_state(dev)
        B, Tn, D = enc.shape
        V, E = emb.shape
        Hl, Hd, A = lm_k.shape[1] // 4, dec_k.shape[1] // 4, q_k.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        enc_flat, r0, r1 = flat_rows(enc)
        assert r1 == 1, "encoder states must be batch-major"
        Tp = r0
        ids = ids[:U].contiguous()
        if samp is not None:
            return AttnDecoderFnV2._forward_sampled(ctx, enc, emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, q_b,
                                                    ap_k, ap_b, out_k, out_b, in_k, in_b, sp_k, sp_b, ids, lens_i32,
                                                    enc_len_i32, U, stash, lm_drop, samp)
        # The LM side of the decoder (embedding -> LM-LSTM -> InputProjection -> decoder-gate pre-activations) reads only
        # the teacher-forced ids and parameters: with the side streams on, it runs on the "dec" stream from the step's
        # start event, concurrently with the encoder, and the main stream joins it here.
        side = _side_stream(dev, "dec")
        main = torch.cuda.current_stream()
        start = _STEP_START.pop(_devkey(dev), None) if stash is not None and stash.get("early_lm") else None
        if side is not None:
            _mark_side(dev, "dec")
            if start is not None:
                side.wait_event(start)
            else:
                side.wait_stream(main)
        ctr = st["ctr_side"] if side is not None else st["ctr"]
        with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
            u = torch.empty((U * B, E), **f32)
            call("e2e_embed_gather", U * B, E, emb, ids, u)
            Wx_lm, Wh_lm, bp_lm = _pack_lstm([lm_k], [lm_b], E, Hl, dev)
            G_lm = gemm(u, Wx_lm, bias=bp_lm)
            hl = torch.zeros((U * B, Hl), **f32)
            C_lm = torch.empty((U * B, Hl), **f32)
            call("e2e_lstm_rec_fwd", B, U, U, Hl, 1, 1, B, G_lm, hl, C_lm, Wh_lm, lens_i32, ctr,
                 ctr.numel() * 4, st["err"], work=float(U), tag="lm_rec_fwd")
            # DropoutWrapper on lm_cell: its OUTPUT is dropped, the recurrent state is not (decoder.py:60-63)
            hl_out = hl
            if lm_drop is not None:
                hl_out = torch.empty_like(hl)
                call("e2e_dropout", hl.numel(), hl, hl_out, float(lm_drop[0]), int(lm_drop[1]), int(lm_drop[2]), 0,
                     seed_dev(lm_drop[1]))
            m = gemm(hl_out, sp_k, bias=sp_b) if sp_k is not None else hl_out
            pre = gemm(m, in_k[:Hd], bias=in_b)                          # [U*B, E]
            # decoder-LSTM kernel in gate-interleaved layout, with the ctx half of InputProjection folded in
            W_ch = torch.empty((D + Hd, 4 * Hd), **f32)
            Wx_dec = torch.empty((E, 4 * Hd), **f32)
            bp_dec = torch.empty((4 * Hd,), **f32)
            call("e2e_lstm_pack_weights", E, Hd, dec_k, dec_b, Wx_dec, 4 * Hd, 0, W_ch[D:], bp_dec)
            gemm(in_k[Hd:], Wx_dec, out=W_ch[:D])                        # W_cx = W_in_c . Wx
            pre_g = gemm(pre, Wx_dec, bias=bp_dec)                       # [U*B, 4Hd]
        if side is not None:
            main.wait_stream(side)
            for t_ in (u, Wx_lm, Wh_lm, bp_lm, G_lm, hl, C_lm, hl_out, m, pre, W_ch, Wx_dec, bp_dec, pre_g):
                t_.record_stream(main)
        HF = gemm(enc_flat, attn_w.view(D, A))
        bufs = dict(cat=torch.empty((U * B, Hd + D), **f32), hprev=torch.zeros((U * B, Hd), **f32),
                    cprev=torch.zeros((U * B, Hd), **f32), acts=torch.empty((U * B, 4 * Hd), **f32),
                    y=torch.empty((U * B, A), **f32), alpha=torch.empty((U * B, Tn), **f32))
        a = _lib.DecPersistArgs()
        a.B, a.U, a.Hd, a.A, a.D, a.Tn, a.Tp = B, U, Hd, A, D, Tn, Tp
        for k, v in dict(W_ch=W_ch, pre_g=pre_g, q_k=q_k, q_b=q_b, attn_v=attn_v, HF=HF, enc=enc_flat,
                         enc_len=enc_len_i32, lens=lens_i32, ctr=st["ctr"], err=st["err"], **bufs).items():
            setattr(a, k, v.data_ptr())
        call("e2e_decoder_persist_fwd", a, work=float(U))
        proj = gemm(bufs["cat"], ap_k, bias=ap_b)
        logits = gemm(proj, out_k, bias=out_b)
        call("e2e_mask_rows", U, B, V, logits, lens_i32)
        AttnDecoderFnV2._save(ctx, (enc, emb, attn_w, attn_v, lm_k, dec_k, q_k, q_b, ap_k, out_k, in_k, sp_k, ids,
                                    lens_i32, enc_len_i32, u, Wx_lm, Wh_lm, G_lm, hl, C_lm, m, pre, pre_g, W_ch, Wx_dec,
                                    HF, proj, bufs["cat"], bufs["hprev"], bufs["cprev"], bufs["acts"], bufs["y"],
                                    bufs["alpha"]),
                              (B, Tn, D, V, E, Hl, Hd, A, U, Tp), stash,
                              (emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, q_b, ap_k, ap_b, out_k, out_b,
                               in_k, in_b, sp_k, sp_b), lm_drop, hl_out)
        return logits

    @staticmethod
    def _save(ctx, tensors, dims, stash, params, lm_drop, hl_out):
        ctx.save_for_backward(*tensors)
        ctx.dims = dims
        ctx.stash = stash
        # flat-gradient-buffer views of the 17 parameters, in the order backward returns their gradients
        ctx.grad_dst = tuple(None if t is None else getattr(t, "grad", None) for t in params)
        ctx.has_sp = params[15] is not None
        ctx.lm_drop = lm_drop
        ctx.hl_out = hl_out if lm_drop is not None else None

    @staticmethod
    def _forward_sampled(ctx, enc, emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, q_b, ap_k, ap_b, out_k, out_b,
                         in_k, in_b, sp_k, sp_b, ids, lens_i32, enc_len_i32, U, stash, lm_drop, samp):
        """Scheduled sampling (attn_decoder.py:130-139, decoder.py:155-180) INSIDE the training forward pass: the loop
        is cut at the steps whose input is sampled.  Between two cuts the inputs are known, so the segment runs exactly
        like the teacher-forced pass -- LM-LSTM over the segment (cell state carried in, h_{a-1} . Wh added to the first
        pre-activations), the batched projections of its rows, the persistent decoder kernel over steps [a, b) -- and
        at a cut the logits of step b-1 alone are formed and `e2e_sample_rows` draws ids[b] from them (the multinomial
        over the unmasked previous logits).  Every buffer the backward pass reads is written once, by the segment that
        owns the step: the backward pass is the ordinary one on the realised ids, and no second forward pass runs.
        samp = (use_sample [U] bools, Philox key, stream offset)."""
        dev = enc.device
        st = _dev_state(dev)
        B, Tn, D = enc.shape
        V, E = emb.shape
        Hl, Hd, A = lm_k.shape[1] // 4, dec_k.shape[1] // 4, q_k.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        enc_flat, Tp, _ = flat_rows(enc)
        use_sample, seed, offset = samp
        ids = ids.clone()
        u = torch.empty((U * B, E), **f32)
        Wx_lm, Wh_lm, bp_lm = _pack_lstm([lm_k], [lm_b], E, Hl, dev)
        G_lm = torch.empty((U * B, 4 * Hl), **f32)
        hl = torch.zeros((U * B, Hl), **f32)
        C_lm = torch.zeros((U * B, Hl), **f32)
        hl_out = torch.empty_like(hl) if lm_drop is not None else hl
        m = torch.empty((U * B, Hd), **f32) if sp_k is not None else hl_out
        pre = torch.empty((U * B, E), **f32)
        pre_g = torch.empty((U * B, 4 * Hd), **f32)
        W_ch = torch.empty((D + Hd, 4 * Hd), **f32)
        Wx_dec = torch.empty((E, 4 * Hd), **f32)
        bp_dec = torch.empty((4 * Hd,), **f32)
        call("e2e_lstm_pack_weights", E, Hd, dec_k, dec_b, Wx_dec, 4 * Hd, 0, W_ch[D:], bp_dec)
        gemm(in_k[Hd:], Wx_dec, out=W_ch[:D])
        HF = gemm(enc_flat, attn_w.view(D, A))
        bufs = dict(cat=torch.empty((U * B, Hd + D), **f32), hprev=torch.zeros((U * B, Hd), **f32),
                    cprev=torch.zeros((U * B, Hd), **f32), acts=torch.empty((U * B, 4 * Hd), **f32),
                    y=torch.empty((U * B, A), **f32), alpha=torch.empty((U * B, Tn), **f32))
        a_ = _lib.DecPersistArgs()
        a_.B, a_.U, a_.Hd, a_.A, a_.D, a_.Tn, a_.Tp = B, U, Hd, A, D, Tn, Tp
        for k, v in dict(W_ch=W_ch, pre_g=pre_g, q_k=q_k, q_b=q_b, attn_v=attn_v, HF=HF, enc=enc_flat,
                         enc_len=enc_len_i32, lens=lens_i32, ctr=st["ctr"], err=st["err"], **bufs).items():
            setattr(a_, k, v.data_ptr())
        cuts = [t for t in range(1, U) if use_sample[t]] + [U]
        a = 0
        for b in cuts:
            rows = slice(a * B, b * B)
            call("e2e_embed_gather", (b - a) * B, E, emb, ids[a:b], u[rows])
            gemm(u[rows], Wx_lm, bias=bp_lm, out=G_lm[rows])
            lens_a = lens_i32
            if a > 0:
                gemm(hl[(a - 1) * B:a * B], Wh_lm[0], out=G_lm[a * B:(a + 1) * B], accumulate=True)
                lens_a = (lens_i32 - a).clamp_(min=0)
                call("e2e_lstm_rec_fwd_carry", B, b - a, b - a, Hl, 1, B, G_lm[rows], hl[rows], C_lm[rows], Wh_lm,
                     lens_a, st["ctr"], st["ctr"].numel() * 4, st["err"], work=float(b - a), tag="lm_rec_fwd")
            else:
                call("e2e_lstm_rec_fwd", B, b, b, Hl, 1, 1, B, G_lm[rows], hl[rows], C_lm[rows], Wh_lm, lens_a,
                     st["ctr"], st["ctr"].numel() * 4, st["err"], work=float(b), tag="lm_rec_fwd")
            if lm_drop is not None:
                call("e2e_dropout", (b - a) * B * Hl, hl[rows], hl_out[rows], float(lm_drop[0]), int(lm_drop[1]),
                     int(lm_drop[2]), a * B * Hl, seed_dev(lm_drop[1]))
            if sp_k is not None:
                gemm(hl_out[rows], sp_k, bias=sp_b, out=m[rows])
            gemm(m[rows], in_k[:Hd], bias=in_b, out=pre[rows])
            gemm(pre[rows], Wx_dec, bias=bp_dec, out=pre_g[rows])
            a_.t0, a_.t1 = a, b
            call("e2e_decoder_persist_fwd", a_, work=float(b - a))
            if b < U:           # ids[b] ~ multinomial(logits of step b-1), rows counted from b*B in the Philox stream
                prj = gemm(bufs["cat"][(b - 1) * B:b * B], ap_k, bias=ap_b)
                lg = gemm(prj, out_k, bias=out_b)
                call("e2e_sample_rows", B, V, lg, V, int(seed), int(offset), b * B, ids[b])
            a = b
        proj = gemm(bufs["cat"], ap_k, bias=ap_b)
        logits = gemm(proj, out_k, bias=out_b)
        call("e2e_mask_rows", U, B, V, logits, lens_i32)
        if stash is not None:
            stash["realized_ids"] = ids
        AttnDecoderFnV2._save(ctx, (enc, emb, attn_w, attn_v, lm_k, dec_k, q_k, q_b, ap_k, out_k, in_k, sp_k, ids,
                                    lens_i32, enc_len_i32, u, Wx_lm, Wh_lm, G_lm, hl, C_lm, m, pre, pre_g, W_ch, Wx_dec,
                                    HF, proj, bufs["cat"], bufs["hprev"], bufs["cprev"], bufs["acts"], bufs["y"],
                                    bufs["alpha"]),
                              (B, Tn, D, V, E, Hl, Hd, A, U, Tp), stash,
                              (emb, attn_w, attn_v, lm_k, lm_b, dec_k, dec_b, q_k, q_b, ap_k, ap_b, out_k, out_b,
                               in_k, in_b, sp_k, sp_b), lm_drop, hl_out)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        (enc, emb, attn_w, attn_v, lm_k, dec_k, q_k, q_b, ap_k, out_k, in_k, sp_k, ids, lens_i32, enc_len_i32, u,
         Wx_lm, Wh_lm, G_lm, hl, C_lm, m, pre, pre_g, W_ch, Wx_dec, HF, proj, cat, hprev, cprev, acts, y,
         alpha) = ctx.saved_tensors
        B, Tn, D, V, E, Hl, Hd, A, U, Tp = ctx.dims
        dev = enc.device
        st = _dev_state(dev)
        f32 = dict(dtype=torch.float32, device=dev)
        dlogits = dlogits.contiguous()
        enc_flat, _, _ = flat_rows(enc)
        nrows = enc_flat.shape[0]
        # ---- critical path (main stream): dlogits -> dcat -> sequential backward -> d enc
        dproj = gemm(dlogits, out_k, tb=True)
        dcat = gemm(dproj, ap_k, tb=True)                             # [U*B, Hd+D]
        out = dict(dcat=dcat, dz=torch.empty((U * B, 4 * Hd), **f32), dch=torch.empty((U * B, D + Hd), **f32),
                   dy=torch.empty((U * B, A), **f32), ds=torch.empty((U * B, Tn), **f32),
                   dc_carry=torch.zeros((B, Hd), **f32))
        denc = torch.zeros((nrows, D), **f32)
        dHF = torch.zeros((nrows, A), **f32)
        dv_part = torch.empty((B * Tn, A), **f32)
        a = _lib.DecPersistArgs()
        a.B, a.U, a.Hd, a.A, a.D, a.Tn, a.Tp = B, U, Hd, A, D, Tn, Tp
        for k, v in dict(W_ch=W_ch, pre_g=pre_g, q_k=q_k, q_b=q_b, attn_v=attn_v, HF=HF, enc=enc_flat,
                         enc_len=enc_len_i32, lens=lens_i32, cat=cat, hprev=hprev, cprev=cprev, acts=acts, y=y,
                         alpha=alpha, ctr=st["ctr"], err=st["err"], **out).items():
            setattr(a, k, v.data_ptr())
        call("e2e_decoder_persist_bwd", a, denc, dHF, dv_part, work=float(U))
        dz, dy = out["dz"], out["dy"]
        gemm(dHF, attn_w.view(D, A), tb=True, out=denc, accumulate=True)
        denc_view = torch.as_strided(denc, (B, Tn, D), (Tp * D, D, 1)) if ctx.needs_input_grad[0] else None

        # ---- parameter gradients (nothing downstream waits for them): 17 tensors in input order
        def param_grads(ctr):
            dout_k = gemm(proj, dlogits, ta=True)
            dout_b = colsum(dlogits)
            dap_k = gemm(cat, dproj, ta=True)
            dap_b = colsum(dproj)
            dq_k = gemm(cat[:, :Hd], dy, ta=True)
            dq_b = colsum(dy)
            dattn_v = colsum(dv_part)
            # gates = pre.Wx + ctx_prev.(W_in_c.Wx) + h_prev.Wh + b   (columns gate-interleaved)
            dW_cx = torch.zeros((D, 4 * Hd), **f32)
            if U > 1:
                gemm(cat[:(U - 1) * B, Hd:], dz[B:], ta=True, out=dW_cx)   # ctx_{t-1} pairs with dz_t
            dWh_dec = gemm(hprev, dz, ta=True).view(1, Hd, 4 * Hd)
            dWx_dec = gemm(pre, dz, ta=True)
            gemm(in_k[Hd:], dW_cx, ta=True, out=dWx_dec, accumulate=True)  # + W_in_c^T . dW_cx
            dbp_dec = colsum(dz)
            ddec_k, ddec_b = _unpack_lstm(dWx_dec, dWh_dec, dbp_dec, E, Hd, 1, dev)
            dpre = gemm(dz, Wx_dec, tb=True)                              # [U*B, E]
            din_k = torch.empty((Hd + D, E), **f32)
            gemm(m, dpre, ta=True, out=din_k[:Hd])
            gemm(dW_cx, Wx_dec, tb=True, out=din_k[Hd:])                  # dW_in_c = dW_cx . Wx^T
            din_b = colsum(dpre)
            dm = gemm(dpre, in_k[:Hd], tb=True)
            dsp_k = dsp_b = None
            if sp_k is not None:
                dsp_k = gemm(hl if ctx.hl_out is None else ctx.hl_out, dm, ta=True)
                dsp_b = colsum(dm)
                dm = gemm(dm, sp_k, tb=True)
            if ctx.lm_drop is not None:
                dmd = torch.empty_like(dm)
                call("e2e_dropout", dm.numel(), dm.contiguous(), dmd, float(ctx.lm_drop[0]), int(ctx.lm_drop[1]),
                     int(ctx.lm_drop[2]), 0, seed_dev(ctx.lm_drop[1]))
                dm = dmd
            call("e2e_lstm_rec_bwd", B, U, U, Hl, 1, 1, B, G_lm, C_lm, Wh_lm, dm, lens_i32, ctr,
                 ctr.numel() * 4, st["err"], work=float(U), tag="lm_rec_bwd")
            dWx_lm = gemm(u, G_lm, ta=True)
            dWh_lm = torch.zeros((1, Hl, 4 * Hl), **f32)
            if U > 1:
                gemm(hl[:(U - 1) * B], G_lm[B:], ta=True, out=dWh_lm[0])
            dbp_lm = colsum(G_lm)
            dlm_k, dlm_b = _unpack_lstm(dWx_lm, dWh_lm, dbp_lm, E, Hl, 1, dev)
            du = gemm(G_lm, Wx_lm, tb=True)
            demb = torch.zeros((V, E), **f32)
            call("e2e_embed_scatter_add", U * B, E, demb, ids, du, E)
            if ctx.stash is not None:
                ctx.stash["emb_values"] = du
            dattn_w = gemm(enc_flat, dHF, ta=True).view(attn_w.shape)
            return [demb, dattn_w, dattn_v, dlm_k, dlm_b, ddec_k, ddec_b, dq_k, dq_b, dap_k, dap_b, dout_k,
                    dout_b, din_k, din_b, dsp_k, dsp_b]

        side = _side_stream(dev, "dec")
        dst = ctx.grad_dst
        need = [True] * 15 + [ctx.has_sp, ctx.has_sp]
        if (side is not None and all((d is not None) or (not n) for d, n in zip(dst, need))
                and all(ctx.needs_input_grad[1 + i] or not need[i] for i in range(17))):
            main = torch.cuda.current_stream()
            _mark_side(dev, "dec")
            side.wait_stream(main)
            with torch.cuda.stream(side):
                grads = param_grads(st["ctr_side"])
                for d, g in zip(dst, grads):
                    if g is not None:
                        d.add_(g.view(d.shape))     # what AccumulateGrad would do, on the side stream
            if _GRAD_READY_HOOK is not None:
                _GRAD_READY_HOOK([d for d in dst if d is not None], side)
            if ctx.stash is not None and ctx.stash.get("emb_values") is not None:
                ctx.stash["emb_values"].record_stream(main)      # read by clip_gradients after the join
            for t_ in list(ctx.saved_tensors) + [dlogits, dproj, dz, dy, dv_part, dHF] + [g for g in grads if g is not None]:
                if t_ is not None:
                    t_.record_stream(side)
            return (denc_view,) + (None,) * 24
        grads = param_grads(st["ctr"])
        return (denc_view,) + tuple(grads) + (None, None, None, None, None, None, None)


# ---------------------------------------------------------------------------
# Losses
# ---------------------------------------------------------------------------

class CrossEntropyFn(torch.autograd.Function):
    """LossUtils.cross_entropy_loss (losses.py:6-35)."""

    @staticmethod
    def forward(ctx, logits, targets, lens_i32):
        U, B = targets.shape
        V = logits.shape[1]
        dev = logits.device
        logits = logits.contiguous()
        targets = targets.contiguous()
        lse = torch.empty((U * B,), dtype=torch.float32, device=dev)
        cost = torch.empty((U * B,), dtype=torch.float32, device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        call("e2e_ce_fwd", U, B, V, logits, targets, lens_i32, lse, cost, loss)
        ctx.save_for_backward(logits, targets, lens_i32, lse)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        logits, targets, lens_i32, lse = ctx.saved_tensors
        U, B = targets.shape
        V = logits.shape[1]
        d = torch.empty_like(logits)
        g = g.contiguous().view(1).to(torch.float32)
        call("e2e_ce_bwd", U, B, V, logits, targets, lens_i32, lse, g, d)
        return d, None, None


class CTCHeadFn(torch.autograd.Function):
    """Auxiliary CTC head on an encoder layer: logits = states.W + b, then
    tf.nn.ctc_loss semantics (blank = C-1), mean over the batch.
    states: [B,T,D] (or its [T,B,D] transpose view) over a padded buffer."""

    @staticmethod
    def forward(ctx, states, kernel, bias, in_lens_i32, labels, label_lens_i32, max_label_len, stash):
        dev = states.device
        f32 = dict(dtype=torch.float32, device=dev)
        flat, r0, r1 = flat_rows(states)
        # batch-major [B,T,D] or the time-major view [T,B,D]?  The strides tell, except when a dimension is 1
        # (a single utterance: [T,1,D] vs [1,T,D]) -- then the number of utterances (lengths) decides.
        nb = in_lens_i32.shape[0]
        if states.shape[0] == nb and states.shape[1] != nb:
            batch_major = True
        elif states.shape[1] == nb and states.shape[0] != nb:
            batch_major = False
        else:
            batch_major = r0 >= r1
        if batch_major:
            B, T = states.shape[0], states.shape[1]
            sb, stt = r0, r1
        else:
            T, B = states.shape[0], states.shape[1]
            sb, stt = r1, r0
        C = kernel.shape[1]
        rows = flat.shape[0]
        logits = gemm(flat, kernel, bias=bias)                         # [rows, C]
        lse = torch.empty((rows,), **f32)
        call("e2e_row_lse", rows, C, logits, C, lse)
        labels = labels.contiguous()
        alpha_ws = torch.empty((_lib.lib().e2e_ctc_workspace_floats(T, B, int(max_label_len)),), **f32)
        loss_b = torch.empty((B,), **f32)
        grad = torch.zeros((rows, C), **f32)
        call("e2e_ctc_fwd_grad", T, B, C, sb, stt, logits, lse, in_lens_i32, labels, labels.stride(0),
             label_lens_i32, max_label_len, alpha_ws, loss_b, grad, 1.0 / B)
        loss = torch.empty((1,), **f32)
        call("e2e_mean", B, loss_b, loss)
        ctx.save_for_backward(states, kernel, grad)
        if stash is not None:
            stash["logits"] = logits
            stash["loss_b"] = loss_b
            stash["layout"] = (B, T, sb, stt)
            consumer = stash.get("consumer_stream")
            if consumer is not None:      # produced on a side stream, consumed on `consumer`
                for t_ in (logits, lse, alpha_ws, loss_b, grad, loss):
                    t_.record_stream(consumer)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        states, kernel, grad = ctx.saved_tensors
        flat, r0, r1 = flat_rows(states)
        g = g.contiguous().view(1).to(torch.float32)
        call("e2e_scale", grad.numel(), grad, g, 1.0)                  # grad *= upstream (in place)
        dk = gemm(flat, grad, ta=True)
        db = colsum(grad)
        dst = None
        if ctx.needs_input_grad[0]:
            dflat = gemm(grad, kernel, tb=True)
            D = states.shape[2]
            dst = torch.as_strided(dflat, states.shape, (r0 * D, r1 * D, 1))
        return dst, dk, db, None, None, None, None, None


def prepare_input(x, Tp, stack=1, stride=1):
    """get_batch frame stacking + initial striding + zero padding (one kernel)."""
    B, T, F = x.shape
    out = torch.empty((B, Tp, F * stack), dtype=torch.float32, device=x.device)
    call("e2e_prepare_input", B, T, F, Tp, stack, stride, x.contiguous(), out)
    return out

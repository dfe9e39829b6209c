"""ctypes binding of libe2e_asr_b200.so (C ABI in include/e2e_asr_b200.h).

There is no CPU fallback: if the shared library is missing (not built) or no
CUDA device is present, every op raises.
"""
import ctypes
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libe2e_asr_b200.so")

# signature codes: p = device/host pointer, i = int, l = long long, z = size_t, f = float
_SIGS = {
    "e2e_gemm": "piiiiiipipipippii",
    "e2e_gemm_lo": "piiiiiippippipippiizzp",
    "e2e_split_rows_f16": "pzippp",
    "e2e_split_lo": "pizpp",
    "e2e_colsum": "piipipi",
    "e2e_lstm_pack_weights": "piipppiipp",
    "e2e_lstm_unpack_grads": "piipppiippi",
    "e2e_lstm_rec_fwd": "piiiiillppppppzp",
    "e2e_lstm_rec_bwd": "piiiiillppppppzp",
    "e2e_lstm_rec_fwd_carry": "piiiillppppppzp",
    "e2e_prepare_input": "piiiiiipp",
    "e2e_embed_gather": "piippp",
    "e2e_embed_scatter_add": "piipppi",
    "e2e_decoder_loop_fwd": "pp",
    "e2e_decoder_loop_bwd": "pp",
    "e2e_decoder_persist_fwd": "pp",
    "e2e_decoder_persist_bwd": "ppppp",
    "e2e_attn_fwd": "piiiiipppppppi",
    "e2e_dec_pointwise_fwd": "piiipppipppippip",
    "e2e_mask_rows": "piiipp",
    "e2e_argmax_rows": "piipip",
    "e2e_ce_fwd": "piiipppppp",
    "e2e_ce_bwd": "piiipppppp",
    "e2e_row_lse": "piipip",
    "e2e_ctc_fwd_grad": "piiillppppipipppf",
    "e2e_sumsq": "pzpppfi",
    "e2e_clip_by_norm": "pzppfpfp",
    "e2e_scale": "pzppf",
    "e2e_mean": "pipp",
    "e2e_axpy": "pzfpp",
    "e2e_adam": "pzppppffff",
    "e2e_dropout": "pzppfQIzp",
    "e2e_lstm_point_fwd": "piipppp",
    "e2e_lstm_point_bwd": "piippppppp",
    "e2e_gru_gate_fwd": "piipppp",
    "e2e_gru_gate_bwd": "piipppppp",
    "e2e_gru_out_fwd": "piipppp",
    "e2e_gru_out_bwd": "piippppppp",
    "e2e_attn_bwd": "piiiiipppppppipppp",
    "e2e_gru_rec_fwd": "piiiiippppppp",
    "e2e_gru_rec_bwd": "piiiiippppppp",
    "e2e_sample_rows": "piipiQIIp",
    "e2e_gemm_f64": "piiipipipip",
    "e2e_gemm_f64d": "piiipipipip",
    "e2e_gemm_f64d_cat": "piiiipipipipippip",
    "e2e_gemm_f64d_lstm": "piiiipipipippippppi",
    "e2e_exp2x_f64": "pzpp",
    "e2e_attn_beam_group_e_f64": "piiiiipppppppi",
    "e2e_lstm_step_f64": "piippppi",
    "e2e_attn_beam_f64": "piiiipppppppi",
    "e2e_attn_beam_group_f64": "piiiiipppppppi",
    "e2e_logsoftmax_topk_f64": "piippdpippp",
    "e2e_embed_gather_f64": "piipppi",
    "e2e_beam_merge": "pp",
    "e2e_beam_gather": "pipp",
}
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_longlong, "z": ctypes.c_size_t, "f": ctypes.c_float,
       "d": ctypes.c_double, "Q": ctypes.c_ulonglong, "I": ctypes.c_uint}


class DecLoopFwdArgs(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ("B", "U", "E", "Hd", "A", "D", "Tn", "Tp", "gemm_mode")] +
                [(n, ctypes.c_void_p) for n in ("in_k", "dec_k", "dec_b", "q_k", "q_b", "attn_v", "pre", "HF",
                                                "enc", "enc_len", "lens", "xh", "cprev", "acts", "cat", "y",
                                                "alpha", "gates_tmp")])


class DecLoopBwdArgs(ctypes.Structure):
    _fields_ = [("f", DecLoopFwdArgs)] + [(n, ctypes.c_void_p) for n in (
        "dcat", "dgates", "dxh", "dy", "dv_part", "dHF", "denc", "dc_carry", "ds")]


class DecPersistArgs(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ("B", "U", "Hd", "A", "D", "Tn", "Tp", "t0", "t1")] +
                [(n, ctypes.c_void_p) for n in ("W_ch", "pre_g", "q_k", "q_b", "attn_v", "HF", "enc", "enc_len",
                                                "lens", "cat", "hprev", "cprev", "acts", "y", "alpha", "dcat", "dz",
                                                "dch", "dy", "ds", "dc_carry", "ctr", "err")])


class BeamMergeArgs(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_int) for n in ("N", "beam", "R", "eos_id")] + [("word_ins_penalty", ctypes.c_double)] +
                [(n, ctypes.c_void_p) for n in ("step", "out_idx", "out_val", "score", "alive", "k_u", "new_tok",
                                                "new_score", "parent", "new_alive", "krow", "par_hist", "tok_hist",
                                                "fin_cnt", "fin_step", "fin_row", "fin_score", "n_live")])


class BeamGatherArgs(ctypes.Structure):
    _fields_ = [("nmat", ctypes.c_int), ("width", ctypes.c_int * 8), ("src", ctypes.c_void_p * 8),
                ("dst", ctypes.c_void_p * 8)]


_lib = None


def lib():
    """Loads the shared library (once).  Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "e2e_asr_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, sig in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes = [_CT[c] for c in sig]
            fn.restype = ctypes.c_int
        l.e2e_last_error.restype = ctypes.c_char_p
        l.e2e_version.restype = ctypes.c_int
        l.e2e_sm_count.restype = ctypes.c_int
        l.e2e_capture_status.restype = ctypes.c_int
        l.e2e_capture_status.argtypes = [ctypes.c_void_p]
        l.e2e_launch_count.restype = ctypes.c_ulonglong
        l.e2e_launch_count.argtypes = [ctypes.c_int]
        l.e2e_set_workspace.restype = ctypes.c_int
        l.e2e_set_workspace.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        l.e2e_set_stream_workspace.restype = ctypes.c_int
        l.e2e_set_stream_workspace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
        l.e2e_set_rec_debug.restype = ctypes.c_int
        l.e2e_set_rec_debug.argtypes = [ctypes.c_void_p]
        l.e2e_set_rec_mode.restype = ctypes.c_int
        l.e2e_set_rec_mode.argtypes = [ctypes.c_int]
        l.e2e_decoder_persist_fits.restype = ctypes.c_int
        l.e2e_decoder_persist_fits.argtypes = [ctypes.c_void_p]
        l.e2e_set_dec_sync.restype = ctypes.c_int
        l.e2e_set_dec_sync.argtypes = [ctypes.c_int]
        l.e2e_ctc_workspace_floats.restype = ctypes.c_size_t
        l.e2e_ctc_workspace_floats.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        l.e2e_lstm_rec_workspace_bytes.restype = ctypes.c_size_t
        l.e2e_lstm_rec_workspace_bytes.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        l.e2e_set_f64_mma.restype = ctypes.c_int
        l.e2e_set_f64_mma.argtypes = [ctypes.c_int]
        l.e2e_set_tc_debug.restype = ctypes.c_int
        l.e2e_set_tc_debug.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
        _lib = l
    return _lib


def exported_symbols():
    return sorted(list(_SIGS.keys()) + ["e2e_last_error", "e2e_version", "e2e_sm_count", "e2e_launch_count", "e2e_capture_status",
                   "e2e_set_workspace", "e2e_set_dec_sync", "e2e_ctc_workspace_floats", "e2e_set_f64_mma", "e2e_lstm_rec_workspace_bytes", "e2e_decoder_persist_fits", "e2e_set_tc_debug", "e2e_set_rec_mode", "e2e_set_rec_debug", "e2e_set_stream_workspace"])


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    return x


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


# E2E_DEBUG_CAPTURE=1: report the first C-ABI call / checkpoint after which a CUDA-graph capture is found invalidated
DEBUG_CAPTURE = bool(os.environ.get("E2E_DEBUG_CAPTURE"))
_capture_reported = False


def capture_checkpoint(where):
    global _capture_reported
    if not DEBUG_CAPTURE or _capture_reported:
        return
    st = lib().e2e_capture_status(stream_ptr())
    if os.environ.get("E2E_DEBUG_CAPTURE") == "2":
        import threading
        sys.stderr.write("[e2e] capture status %d stream %x thread %s after %s\n"
                         % (st, stream_ptr(), threading.current_thread().name, where))
    if st == 2 or st == -1:
        _capture_reported = True
        import traceback
        sys.stderr.write("[e2e] capture found INVALIDATED (status %d) after: %s\n" % (st, where))
        traceback.print_stack(limit=8, file=sys.stderr)


class Profiler(object):
    """Optional per-entry-point device timing: CUDA events recorded on the launching
    stream around every C call (bench.py turns this on for the timed region)."""

    def __init__(self):
        self.records = []      # (name, start_event, end_event, work, stream)

    def summary(self, main_stream=None):
        """Per entry point: total event time, calls, algorithmic work; `main_ms` / `main_calls` / `main_work` count
        only the calls enqueued on `main_stream` (the critical path: side-stream events include waiting for SMs)."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, work, stream in self.records:
            d = out.setdefault(name, dict(ms=0.0, calls=0, work=0.0, main_ms=0.0, main_calls=0, main_work=0.0))
            t = e0.elapsed_time(e1)
            d["ms"] += t
            d["calls"] += 1
            d["work"] += work
            if main_stream is not None and stream == main_stream:
                d["main_ms"] += t
                d["main_calls"] += 1
                d["main_work"] += work
        return out


PROFILER = None


def launch_count(reset=False):
    return int(lib().e2e_launch_count(1 if reset else 0))


def call(name, *args, work=0.0, tag=None):
    """Invoke `name(stream, *args)`; tensors are passed as raw device pointers."""
    l = lib()
    if not torch.cuda.is_available():
        raise RuntimeError("e2e_asr_b200: no CUDA device (there is no CPU fallback)")
    conv = []
    for a in args:
        if isinstance(a, ctypes.Structure):
            conv.append(ctypes.addressof(a))
        else:
            conv.append(_ptr(a))
    prof = PROFILER
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(l, name)(stream_ptr(), *conv)
    if prof is not None:
        e1.record()
        prof.records.append((tag or name, e0, e1, work, stream_ptr()))
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, l.e2e_last_error().decode()))
    if DEBUG_CAPTURE:
        capture_checkpoint(tag or name)

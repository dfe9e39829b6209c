#!/usr/bin/env python
"""Benchmark of the hot path: teacher-forced fwd+bwd training step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2]

Prints ONE JSON line (rank 0).  metric = train frames/sec (valid log-mel frames,
sum of logmel_len) for fwd + bwd + gradient clipping on BASELINE.json configs[1]
("cfg2": B=64/GPU, T=700, F=120, 4-layer pyramidal BiLSTM 256/dir, char attention
decoder, phone+state aux CTC), weak scaling (per-GPU batch fixed).

  value : inputs already resident in HBM when the timed region starts
  e2e   : the same step through the public API from pinned HOST buffers (H2D of
          the batch + D2H of the loss inside the timed region)
  --impl reference : the reference's CPU path.  TensorFlow-1.x / Python 2 cannot run
          in this image, so this is the line-by-line NumPy restatement in oracle/
          ("port"), float32, all host cores, on a bounded sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_frames_per_sec_fwd_bwd"
UNIT = "frames/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_step(cfg_name, sample_B, steps, warmup, threads=None):
    """Times the CPU restatement (oracle/) on a bounded sample: the first sample_B
    utterances' worth of the workload (same T/F/H/U, batch reduced)."""
    from e2e_asr_b200 import synth
    from oracle import model as om
    cfg = synth.get_config(cfg_name, B=sample_B)
    w = synth.make_weights(cfg)
    batch = synth.make_batch(cfg)
    frames = int(batch["logmel_len"].sum())
    kw = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, dtype=np.float32)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        om.train_step(w, batch, **kw)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return frames / sec, sec, frames, cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_B = 8
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    fps, sec, frames, cfg = cpu_reference_step(args.config, sample_B, steps, warmup)
    sample = ("%s shapes with batch %d of %d utterances (%d valid frames/step), float32 NumPy restatement "
              "of the reference graph (oracle/model.py), %d step(s) after %d warm-up"
              % (args.config, sample_B, cfg_B(args.config), frames, steps, warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, 1, note="CPU arm runs a bounded sample"),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cfg_B(name):
    from e2e_asr_b200 import synth
    return synth.CONFIGS[name]["B"]


def workload_config(name, n_gpus, note=None):
    from e2e_asr_b200 import synth
    c = synth.CONFIGS[name]
    d = {"workload": "BASELINE.json configs[1] (%s): Switchboard-300h-shaped teacher-forced fwd+bwd+clip" % name,
         "batch_per_gpu": c["B"], "global_batch": c["B"] * n_gpus, "frames_T": c["T"], "feat_F": c["F"],
         "hidden_per_dir": c["H"], "enc_layers": c["L"], "vocab": c["V"], "target_len_U": c["U"],
         "aux_ctc": {k: {"depth": v[0], "vocab": v[1]} for k, v in c["ctc"].items()},
         "parallelism": "dp%d" % n_gpus, "dropout": "off (out_prob=1)", "samp_prob": 0.0,
         "l2": "activations per step (~2 GB) exceed the 126 MB L2; no explicit flush"}
    if note:
        d["note"] = note
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--gemm", default="tf32x3", help="fp32 | tf32x3 | bf16x2 | bf16 (default: tf32x3, the fp32-accurate tensor-core mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true", help="skip the beam-decode utt/s side measurement")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of its CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from e2e_asr_b200 import _lib, ops, synth
    from e2e_asr_b200 import dist as edist
    from e2e_asr_b200.testing import build_model

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    # NCCL prints its version banner on stdout at communicator creation: keep stdout for the ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, world, local = edist.init_from_env()
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        if world > 1:
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if args.gemm:
        ops.set_gemm_mode(args.gemm)
    W = max(args.warmup, 3)
    K = args.steps
    peaks = load_peaks()

    cfg = synth.get_config(args.config)
    weights = synth.make_weights(cfg)
    batch = synth.make_batch(cfg, seed=synth.DATA_SEED + rank)      # weak scaling: every rank its own batch
    frames = int(batch["logmel_len"].sum())
    reducer = edist.GradAllReducer() if world > 1 else None
    model = build_model(cfg, weights, device=dev, reducer=reducer)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- eager pass: per-kernel CUDA events (breakdown, roofline) + host enqueue cost
    prepared = model.get_batch(batch)
    for _ in range(W):
        model.run_step(prepared=prepared)
    # untimed: let the caching allocator reach its steady state for the asynchronous loop (no cudaMalloc inside
    # the timed region); at most 8 extra steps
    # (the exit is a COLLECTIVE decision: every step holds a gradient all-reduce, so all ranks must run the same
    # number of steps -- ranks see different batches and settle at different times)
    for _ in range(8):
        n0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
        model.run_step(prepared=prepared)
        model.run_step(prepared=prepared)
        grew = float(torch.cuda.memory_stats().get("num_device_alloc", 0) != n0)
        if max_over_ranks(grew) == 0.0:
            break
    barrier()
    ops.check_device_errors(dev)
    use_graph = not args.no_graph
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_loop(step_fn, profile):
        """K steps between barriers; returns (ms max over ranks, host enqueue ms/step, launches, clocks)."""
        prof = _lib.Profiler() if profile else None
        _lib.PROFILER = prof
        _lib.launch_count(reset=True)
        sampler = ClockSampler(local) if rank == 0 else None
        barrier()
        e0.record()
        model.host_wait_s = 0.0
        h0 = time.perf_counter()
        for _ in range(K):
            step_fn()
        host_s = time.perf_counter() - h0
        e1.record()
        barrier()
        t_ms = max_over_ranks(e0.elapsed_time(e1))
        host_ms = max_over_ranks((host_s - model.host_wait_s) * 1e3 / K)   # enqueue work, excluding waits on the GPU
        clk = sampler.stop() if sampler else None
        n = _lib.launch_count()
        _lib.PROFILER = None
        return t_ms, host_ms, n, clk, prof

    ops.TAG_GEMM_SHAPES = True          # one profiler entry per GEMM shape (folded into a class below)
    ms_eager, host_eager_ms, launches, clocks, prof = timed_loop(lambda: model.run_step(prepared=prepared), True)
    ops.TAG_GEMM_SHAPES = False
    ms, host_busy_ms = ms_eager, host_eager_ms
    step_host = lambda: model.run_step(batch)
    mode = "eager (~200 launches per step from Python)"
    gs = None
    if use_graph:
        # ---------------- device-resident inputs ("value"): the step replayed from its CUDA graph
        # (a capture failure must not cost the run its number: every rank then times the eager step instead)
        try:
            gs = model.graphed_step(batch)
            failed = 0.0
        except Exception as e:          # noqa: BLE001
            sys.stderr.write("bench.py: CUDA-graph capture failed (%r); timing the eager step\n" % (e,))
            mode = "eager (graph capture failed: %s)" % (str(e)[:120],)
            failed = 1.0
        if max_over_ranks(failed) > 0.0:
            gs, use_graph = None, False
    if use_graph:
        for _ in range(W):
            gs.step()
        ms, host_busy_ms, n_tail, clocks, _ = timed_loop(gs.step, False)
        launches = n_tail + gs.launches_per_step * K         # kernels in the replayed graph + the eager tail
        step_host = lambda: gs.step(batch)
        mode = "CUDA graph replay of fwd+bwd (%d kernels) + eager all-reduce/clip" % gs.launches_per_step
    loss_val = float(model.total_loss)
    total_frames = sum_over_ranks(frames)
    value = total_frames * K / (ms * 1e-3)

    # ---------------- end to end from host buffers ("e2e"): the batch lives in PINNED host memory (what a loader with
    # pinned output buffers hands over); every step copies it host -> device and reads the loss back
    host_batch = dict(batch)
    host_batch["logmel"] = torch.from_numpy(np.ascontiguousarray(batch["logmel"], np.float32)).pin_memory()
    if use_graph:
        # the copy of the NEXT step's inputs is started before the loss of this step is read back, so it overlaps
        # the running step (double-buffered input pipeline); every step still has its own H2D inside the region
        def step_host():
            gs.step()
            gs.prefetch(host_batch)
        gs.prefetch(host_batch)
    else:
        step_host = lambda: model.run_step(host_batch)
    for _ in range(2):
        step_host()
        float(model.total_loss)
    barrier()
    e0.record()
    for _ in range(K):
        step_host()                             # H2D of a batch from pinned host memory inside
        _ = float(model.total_loss)             # D2H of the step's result
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    h2d = int(batch["logmel"].nbytes + sum(batch[k].nbytes for k in batch if k != "logmel" and k != "utt_id"
                                            and hasattr(batch[k], "nbytes")))
    e2e = {"value": total_frames * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K}
    ops.check_device_errors(dev)

    if world > 1:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    # ---------------- per-kernel-class breakdown and roofline (CUDA events of the eager timed pass)
    raw = prof.summary(main_stream=torch.cuda.current_stream().cuda_stream)
    # GEMM calls were tagged per shape: fold them into one class for the breakdown, keep the shapes for the roofline
    shapes = {k: v for k, v in raw.items() if k.startswith("gemm M=")}
    summ = {k: v for k, v in raw.items() if k not in shapes}
    if shapes:
        summ["e2e_gemm"] = {"ms": sum(v["ms"] for v in shapes.values()), "calls": sum(v["calls"] for v in shapes.values()),
                            "work": sum(v["work"] for v in shapes.values())}
    tot = sum(v["ms"] for v in summ.values())
    breakdown = {k: {"ms_per_step": v["ms"] / K, "calls_per_step": v["calls"] / K,
                     "share": v["ms"] / tot if tot else 0.0} for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}
    top = max(summ.items(), key=lambda kv: kv[1]["ms"])
    name, tv = top
    if name == "e2e_gemm":
        peak = peaks["tc_sustained"]
        gm = ops.get_gemm_mode()
        nprod = {0: 1, 1: 3, 2: 1, 3: 3}[gm]
        # tensor-pipe cost of one algorithmic product in bf16-MMA units: 3xTF32 issues 3 kind::tf32 MMAs and tf32 runs
        # at half the bf16 rate (6 units); bf16x2 issues 3 kind::f16 MMAs (3 units).  "achieved"/"frac" stay
        # ALGORITHMIC FLOPs against the bf16 peak as the contract asks; the mode ceiling is peak / units.
        units = {0: 1, 1: 6, 2: 1, 3: 3}[gm]
        # the dominant KERNEL = the GEMM shape with the largest time ON THE MAIN STREAM (the critical path; one launch
        # configuration of gemm_tc_kernel).  Event durations of side-stream GEMMs include the time they wait for SMs
        # held by the persistent decoder / recurrence kernels, so they say little about the kernel; the whole class
        # (all shapes, all streams) is reported next to it.
        on_main = {k: v for k, v in shapes.items() if v["main_calls"] > 0} or shapes
        sname, sv = max(on_main.items(), key=lambda kv: kv[1]["main_ms"] if kv[1]["main_calls"] else kv[1]["ms"])
        if sv["main_calls"]:
            sv = {"ms": sv["main_ms"], "calls": sv["main_calls"], "work": sv["main_work"]}
        ach = sv["work"] / (sv["ms"] * 1e-3) / 1e12
        ach_class = tv["work"] / (tv["ms"] * 1e-3) / 1e12
        roof = {"kernel": "gemm_tc_kernel<%d> %s (the GEMM shape with the largest time on the main stream; incl. its "
                          "operand split passes if any)" % (gm, sname),
                "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": peaks["source"] + ", sustained bf16",
                "algorithmic_gflop_per_launch": sv["work"] / sv["calls"] / 1e9,
                "avg_launch_ms": sv["ms"] / sv["calls"], "launches_per_step": sv["calls"] / K,
                "mma_products_per_flop": nprod,
                "mode_ceiling_tflops": peak / units,
                "frac_of_mode_ceiling": ach / (peak / units),
                "class_aggregate": {"kernel": "all e2e_gemm calls (every dense projection incl. dX/dW, %d shapes)"
                                              % len(shapes),
                                    "algorithmic_gflop_per_step": tv["work"] / K / 1e9,
                                    "achieved": ach_class, "frac": ach_class / peak,
                                    "frac_of_mode_ceiling": ach_class / (peak / units),
                                    "avg_launch_ms": tv["ms"] / tv["calls"]},
                "top_shapes": [{"shape": k, "ms_per_step": v["ms"] / K, "on_main_stream": v["main_calls"] > 0,
                                "tflops": v["work"] / (v["ms"] * 1e-3) / 1e12}
                               for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:6]],
                "note": "events of GEMMs on the weight-gradient / CTC side streams overlap the main stream and wait "
                        "for SMs held by the persistent kernels: the class total exceeds its share of the step"}
    else:
        # recurrence / decoder loop: latency bound; algorithmic HBM bytes are small, report per-timestep latency
        steps_total = tv["work"]
        roof = {"kernel": name, "bound": "hbm", "achieved": None, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": None, "traffic": None, "peak_source": peaks["source"],
                "us_per_timestep": tv["ms"] * 1e3 / max(steps_total, 1.0), "note": "latency-bound sequential kernel"}
    rec = {k: summ[k] for k in summ if "rec" in k}
    extra = {k: {"us_per_timestep": v["ms"] * 1e3 / max(v["work"], 1.0)} for k, v in rec.items()}
    for k in ("e2e_decoder_loop_fwd", "e2e_decoder_loop_bwd", "e2e_decoder_persist_fwd", "e2e_decoder_persist_bwd"):
        if k in summ:
            extra[k] = {"us_per_timestep": summ[k]["ms"] * 1e3 / max(summ[k]["work"], 1.0)}
    memory_bound = {}
    try:
        # memory-bound kernels: ALGORITHMIC bytes per call / event time, against the measured HBM peak
        D_, T_enc = 2 * cfg.H, synth.pyramid_lens([cfg.T], synth.depth_reductions(cfg, cfg.L))[0]
        rows_ce = cfg.U * cfg.B * cfg.V * 4
        mem_bytes = {"e2e_ce_fwd": rows_ce,                         # reads the logits once (lse, picked cost)
                     "e2e_ce_bwd": 2 * rows_ce,                     # reads logits, writes d logits
                     # attention(): HF [B,T_enc,A] + enc [B,T_enc,D] read once per decoder step (14.4 MB at cfg-2; the
                     # working set fits the 126 MB L2, so this is L2 -> SM traffic, not HBM)
                     "e2e_decoder_persist_fwd": cfg.U * 4 * cfg.B * int(T_enc) * (cfg.A + D_)}
        ctc_bytes = 0
        for t_, (depth, vocab) in cfg.ctc.items():
            Td = int(synth.pyramid_lens([cfg.T], synth.depth_reductions(cfg, depth))[0])
            ctc_bytes += 2 * 4 * Td * cfg.B * (vocab + 1)           # read the logits, write their gradient
        n_heads = max(len(cfg.ctc), 1)
        mem_bytes["e2e_ctc_fwd_grad"] = ctc_bytes / n_heads         # per call (one call per head)
        for k, bytes_per_call in mem_bytes.items():
            if k in summ and summ[k]["ms"] > 0:
                gbs = bytes_per_call * summ[k]["calls"] / (summ[k]["ms"] * 1e-3) / 1e9
                memory_bound[k] = {"algorithmic_mb_per_call": bytes_per_call / 1e6, "achieved_gbs": gbs,
                                   "frac_of_hbm_peak": gbs / peaks["hbm"]}
        if "e2e_ctc_fwd_grad" in memory_bound:
            memory_bound["e2e_ctc_fwd_grad"]["note"] = ("alpha/beta sweep = T_l sequential frames per utterance (one warp "
                                                        "each): latency-bound, on its own stream beside the decoder")
        if "e2e_decoder_persist_fwd" in memory_bound:
            memory_bound["e2e_decoder_persist_fwd"]["note"] = ("attention operands only; L2-resident; the kernel also "
                                                               "runs the gate GEMMs and 3 grid barriers per step")

    except Exception as e:          # supplementary section: never cost the run its JSON line
        memory_bound = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {0: "f32", 1: "tf32x3", 2: "bf16", 3: "bf16x2"}[ops.get_gemm_mode()], "data": "synthetic",
        "config": workload_config(args.config, world), "frames_per_step_per_gpu": frames,
        "padded_frames_per_step_per_gpu": cfg.B * cfg.T, "loss": loss_val,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launches_per_step": launches / K,
        "step_mode": mode, "host_enqueue_ms_per_step": host_busy_ms,
        "eager": {"ms_per_step": ms_eager / K, "host_enqueue_ms_per_step": host_eager_ms,
                  "note": "same K steps launched kernel by kernel; the per-kernel events of breakdown/roofline come "
                          "from this pass (a graph replay runs the identical kernels)"},
        "roofline": roof, "breakdown": breakdown, "sequential_kernels": extra, "memory_bound_kernels": memory_bound,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_B = 8
        fps, sec, fr, _ = cpu_reference_step(args.config, sample_B, 2, 1)
        line["cpu_baseline"] = {
            "value": fps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%s shapes, batch %d of %d utterances (%d valid frames), float32 NumPy restatement "
                      "(oracle/model.py), mean of 2 steps after 1 warm-up, %.1f s/step"
                      % (args.config, sample_B, cfg.B, fr, sec)}
    if world == 1 and not args.no_beam and args.config == "cfg2":
        line["beam_decode"] = beam_decode_rate(args.config, dev, cpu=not args.no_cpu_baseline)
    print(json.dumps(line))
    sys.stdout.flush()


def beam_decode_rate(cfg_name, dev, n_utts=256, beam=10, cpu=True):
    """Second half of BASELINE.json's metric: beam-search decoding (BASELINE configs[2]: beam width 10, a synthetic
    eval batch of 256 utterances with T_enc in [50, 88]) in utterances/s through BeamSearch.decode_batch, next to the
    CPU restatement of beam_search.py (serial over utterances like eval_model.py:194-195) on ONE utterance."""
    import torch
    from e2e_asr_b200 import synth
    from e2e_asr_b200.beam_search import BeamSearch
    cfg = synth.get_config(cfg_name)
    w = synth.make_weights(cfg)
    rng = np.random.Generator(np.random.PCG64(17))
    encs = [(np.tanh(rng.standard_normal((int(rng.integers(50, 89)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(n_utts)]
    sp = BeamSearch.class_params()
    sp.beam_size = beam
    bs = BeamSearch(w, sp, device=dev)
    bs.decode_batch(encs[:8])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = bs.decode_batch(encs)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = {"value": n_utts / dt, "unit": "utt/s", "beam_size": beam, "n_utts": n_utts,
           "mean_output_len": float(np.mean([len(o) for o in out])),
           "note": "random-init weights: hypotheses run to the 120-step limit (worst case); wall clock incl. the "
                   "host-side k^2 candidate merge"}
    if cpu:
        from oracle import beam as ob
        t0 = time.perf_counter()
        ref = ob.beam_search(w, encs[0], beam_size=beam)
        dt1 = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 1.0 / dt1, "unit": "utt/s", "cores": 1, "kind": "port",
                               "sample": "1 utterance, Python restatement of beam_search.py (oracle/beam.py)",
                               "ids_equal": bool(np.array_equal(ref, out[0]))}
    return res


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the hot path: teacher-forced fwd+bwd training step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg4|cfg5]
                  [--scaling weak|strong] [--defaults]

Prints ONE JSON line (rank 0).  metric = train frames/sec (valid log-mel frames,
sum of logmel_len) for fwd + bwd + gradient all-reduce + clipping on BASELINE.json
configs[1] ("cfg2": B=64/GPU, T=700, F=120, 4-layer pyramidal BiLSTM 256/dir, char
attention decoder, phone+state aux CTC).  --scaling weak (default): the config's batch
per GPU; strong: the config's batch split over the GPUs.

  value            inputs already resident in HBM when the timed region starts
  e2e              the same step through the public API from pinned HOST buffers (H2D of
                   the batch + D2H of the loss inside the timed region)
  roofline         the DOMINANT kernel = the kernel with the largest time on the main
                   stream (the critical path): algorithmic rate vs the measured peak, its
                   latency floor for a sequential kernel, DRAM traffic from profiles/
  pct_of_roofline  SURVEY.md 8(d): sum_k (algorithmic time of kernel k at its bound) / step
  cpu_baseline     the NumPy restatement (oracle/) on a bounded sample of the workload
  --impl reference the reference's CPU path on the FULL batch: TensorFlow-1.x / Python 2
                   cannot run in this image, so this is the line-by-line NumPy restatement in
                   oracle/ ("port"), float32, all host cores (median of 5), plus a
                   single-thread figure (the reference's own setting, train.py:178)
  --defaults       the reference's default training settings (out_prob = out_prob_dec = 0.9
                   captured in the graph; samp_prob = 0.1 runs the eager step)
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
    """Times the CPU restatement (oracle/) on `sample_B` utterances of the workload (same T/F/H/U; sample_B = the
    config's batch size is the whole step).  threads = 1 is what the reference itself configures
    (intra_op_parallelism_threads=1, train.py:178); None = all host cores.  Returns the MEDIAN step."""
    from threadpoolctl import threadpool_limits
    from e2e_asr_b200 import synth
    from oracle import model as om
    cfg = synth.get_config(cfg_name, B=sample_B)
    w = synth.make_weights(cfg)
    batch = synth.make_batch(cfg)
    frames = int(batch["logmel_len"].sum())
    kw = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, dtype=np.float32)
    times = []
    with threadpool_limits(limits=threads):
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            om.train_step(w, batch, **kw)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    sec = float(np.median(times))
    return frames / sec, sec, frames, cfg


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on the box's host cores.  TensorFlow-1.x /
    Python 2 cannot run in this image, so it is the NumPy restatement (oracle/model.py, "port"), float32, on the FULL
    batch of the workload with all host cores: 2 warm-up steps + the median of 5, bounded by
    --steps / --warmup from above; plus one single-thread figure (the reference's own train.py:178 setting) on an
    8-utterance sample so that the whole arm still ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    full_B = cfg_B(args.config)
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 2))
    fps, sec, frames, cfg = cpu_reference_step(args.config, full_B, steps, warmup)
    sample = ("%s at FULL batch %d (%d valid frames/step), float32 NumPy restatement of the reference graph "
              "(oracle/model.py; restated reference, not TF), all %d host cores, median of %d step(s) after %d warm-up, "
              "%.1f s/step" % (args.config, full_B, frames, cores, steps, warmup, sec))
    one = None
    try:
        sB = min(8, full_B)
        f1, s1, fr1, _ = cpu_reference_step(args.config, sB, 1, 0, threads=1)
        one = {"value": f1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d of %d utterances (%d valid frames), ONE thread (the reference's own "
                         "intra_op_parallelism_threads=1, train.py:178), 1 step, %.1f s" % (sB, full_B, fr1, s1)}
    except Exception as e:          # noqa: BLE001
        one = {"error": repr(e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, 1, note="CPU arm: the same full batch on the host cores"),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_thread": one},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cfg_B(name):
    from e2e_asr_b200 import synth
    return synth.CONFIGS[name]["B"]


WORKLOADS = {"cfg1": "BASELINE.json configs[0] (cfg1): base_params defaults, batch 4 x ~200 frames",
             "cfg2": "BASELINE.json configs[1] (cfg2): Switchboard-300h-shaped teacher-forced fwd+bwd+clip",
             "cfg4": "BASELINE.json configs[3] (cfg4): long-utterance stress, batch 32 x 2000 frames",
             "cfg5": "BASELINE.json configs[4] (cfg5): wide encoder, 5-layer BiLSTM 512/dir, batch 256"}


def workload_config(name, n_gpus, note=None, batch_per_gpu=None, settings=None):
    from e2e_asr_b200 import synth
    c = synth.CONFIGS[name]
    bpg = c["B"] if batch_per_gpu is None else batch_per_gpu
    d = {"workload": WORKLOADS.get(name, name),
         "batch_per_gpu": bpg, "global_batch": bpg * n_gpus, "frames_T": c["T"], "feat_F": c["F"],
         "hidden_per_dir": c["H"], "enc_layers": c["L"], "vocab": c["V"], "target_len_U": c["U"],
         "aux_ctc": {k: {"depth": v[0], "vocab": v[1]} for k, v in c["ctc"].items()},
         "parallelism": "dp%d" % n_gpus, "dropout": "off (out_prob=1)", "samp_prob": 0.0,
         "l2": "activations per step (~2 GB) exceed the 126 MB L2; no explicit flush"}
    if settings:
        d.update(settings)
    if note:
        d["note"] = note
    return d



# Latency floors of the sequential kernels, in SM cycles per timestep, from the micro-benchmarks behind DESIGN.md 4.1 /
# 4.2 (scratch/ubench.cu and in-kernel clock64 stamps):
#   recurrence, two interleaved slices per cluster (NS = 2, the encoder layers at B = 64): the MMA warps must issue
#     2 x 1024 cycles of mma.sync per timestep (64 MMAs per warp and slice at 8 cycles/SMSP, 2 warps per SMSP); the
#     h_t exchange (L2 tile + one multicast: 910 cycles; backward: bulk DSMEM reduce-scatter, 1416) hides behind the
#     other slice -> max(2048, 1416) = 2048
#   recurrence, one slice per cluster (NS = 1, the LM-LSTM): k-loop 1024 + exchange 910 back to back = 1934
#   decoder loop: three grid-wide barriers per step, 2600 cycles each (128 co-resident CTAs, one L2 atomic + poll)
# latency floors in cycles per timestep: fp16 split scheme = 768 cycles of MMA issue per slice-step
# (3 x m16n8k16 per k16 and n-tile at one MMA per 2 cycles per SM); encoder layers interleave two slices per cluster
# (2 x 768, the exchange hides behind the other slice), the LM-LSTM runs one slice per cluster (768 + the exchange:
# 910-cycle multicast forward, 1416 cycles of bulk DSMEM copies backward)
FLOOR_CYCLES = {"enc_rec_fwd": 1536, "enc_rec_bwd": 1536, "lm_rec_fwd": 1678, "lm_rec_bwd": 2184,
                "e2e_decoder_persist_fwd": 7800, "e2e_decoder_persist_bwd": 7800}
KERNEL_OF = {"enc_rec_fwd": "rec_fwd_ws_kernel<16,2,8>", "enc_rec_bwd": "rec_bwd_ws_kernel<16,2,8>",
             "lm_rec_fwd": "rec_fwd_ws_kernel<16,1,4>", "lm_rec_bwd": "rec_bwd_ws_kernel<16,1,4>",
             "e2e_decoder_persist_fwd": "dec_fwd_persist_kernel", "e2e_decoder_persist_bwd": "dec_bwd_persist_kernel"}


def load_traffic():
    """Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the profiled kernels, from the
    `ncu --set full` captures summarised under profiles/ (never measured inside a bench run)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def roofline_report(raw, cfg, peaks, K, ms_per_step, clocks, gemm_mode):
    """Folds the per-call CUDA events of the eager pass into: the per-class breakdown, the `roofline` object of the
    DOMINANT kernel (largest time on the main stream = the critical path), the latency-floor fractions of the four
    sequential kernels, the memory-bound kernels' achieved GB/s, and SURVEY.md 8(d)'s end-to-end figure
    pct_of_roofline = sum_k (algorithmic time of kernel k at its bound) / measured step time."""
    from e2e_asr_b200 import synth
    shapes = {k: v for k, v in raw.items() if k.startswith("gemm M=")}
    summ = {k: v for k, v in raw.items() if k not in shapes}
    if shapes:
        summ["e2e_gemm"] = {k2: sum(v[k2] for v in shapes.values()) for k2 in
                            ("ms", "calls", "work", "main_ms", "main_calls", "main_work")}
    tot = sum(v["ms"] for v in summ.values())
    breakdown = {k: {"ms_per_step": v["ms"] / K, "calls_per_step": v["calls"] / K,
                     "main_stream_ms_per_step": v["main_ms"] / K, "share": v["ms"] / tot if tot else 0.0}
                 for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}
    sm_mhz = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
    peak_tc = peaks["tc_sustained"]
    # tensor-pipe cost of one algorithmic product: 3xTF32 = 3 MMAs ("divides the peak by its split count")
    # which moreover run at half the bf16 rate (kind::tf32) -> issue ceiling peak / 6; bf16x2 = 3 bf16 MMAs
    split = {0: 1, 1: 3, 2: 1, 3: 3, 4: 3}[gemm_mode]
    units = {0: 1, 1: 6, 2: 1, 3: 3, 4: 3}[gemm_mode]
    traffic = load_traffic()
    H, Bn = cfg.H, cfg.B
    nd = 2 if cfg.get("bi_dir", True) else 1

    # ---- sequential kernels: measured us / timestep against the latency floor, and their tensor-pipe rate
    flop_per_step = {"enc_rec_fwd": 2.0 * Bn * H * 4 * H * nd, "enc_rec_bwd": 2.0 * Bn * H * 4 * H * nd,
                     "lm_rec_fwd": 2.0 * Bn * cfg.Hl * 4 * cfg.Hl, "lm_rec_bwd": 2.0 * Bn * cfg.Hl * 4 * cfg.Hl,
                     # gates [B, D+Hd] x [D+Hd, 4Hd] + query [B, Hd] x [Hd, A]; backward: the two transposed products
                     "e2e_decoder_persist_fwd": 2.0 * Bn * ((2 * H + cfg.Hd) * 4 * cfg.Hd + cfg.Hd * cfg.A),
                     "e2e_decoder_persist_bwd": 2.0 * Bn * ((2 * H + cfg.Hd) * 4 * cfg.Hd + cfg.Hd * cfg.A)}
    seq, floor_ms = {}, 0.0
    for k, v in summ.items():
        if not ("rec" in k or k.startswith("e2e_decoder")):
            continue
        steps = max(v["work"], 1.0)
        us = v["ms"] * 1e3 / steps
        d = {"kernel": KERNEL_OF.get(k, k), "us_per_timestep": us, "timesteps_per_step": steps / K,
             "ms_per_step": v["ms"] / K, "share_of_step": v["ms"] / K / ms_per_step}
        if k in FLOOR_CYCLES and (H == 256 or not k.startswith("enc_rec")):
            fl_us = FLOOR_CYCLES[k] / sm_mhz
            d.update(floor_cycles=FLOOR_CYCLES[k], floor_us=fl_us, frac_of_latency_floor=fl_us / us)
            floor_ms += fl_us * steps / K * 1e-3
        if k in flop_per_step:
            d["tensor_tflops"] = flop_per_step[k] * steps / (v["ms"] * 1e-3) / 1e12
            d["frac_of_tensor_peak"] = d["tensor_tflops"] / peak_tc
        seq[k] = d

    # ---- memory-bound kernels: ALGORITHMIC bytes per call / event time vs the measured HBM peak
    memory_bound, mem_ms = {}, 0.0
    D_, T_enc = 2 * H, int(synth.pyramid_lens([cfg.T], synth.depth_reductions(cfg, cfg.L))[0])
    rows_ce = cfg.U * Bn * cfg.V * 4
    attn_fwd = cfg.U * 4.0 * Bn * T_enc * (cfg.A + D_)
    mem_bytes = {"e2e_ce_fwd": rows_ce, "e2e_ce_bwd": 2 * rows_ce,
                 # attention(): HF [B,T_enc,A] + enc [B,T_enc,D] read once per decoder step; backward re-reads both and
                 # read-modify-writes dHF / dEnc = 3x (the working set fits the 126 MB L2: L2 -> SM traffic, not HBM)
                 "e2e_decoder_persist_fwd": attn_fwd, "e2e_decoder_persist_bwd": 3.0 * attn_fwd}
    ctc_bytes = []
    for t_, (depth, vocab) in cfg.ctc.items():
        Td = int(synth.pyramid_lens([cfg.T], synth.depth_reductions(cfg, depth))[0])
        ctc_bytes.append(2.0 * 4 * Td * Bn * (vocab + 1))           # read the logits, write their gradient
    if ctc_bytes:
        mem_bytes["e2e_ctc_fwd_grad"] = sum(ctc_bytes) / len(ctc_bytes)     # per call (one call per head)
    for k, bytes_per_call in mem_bytes.items():
        if k in summ and summ[k]["ms"] > 0:
            gbs = bytes_per_call * summ[k]["calls"] / (summ[k]["ms"] * 1e-3) / 1e9
            memory_bound[k] = {"algorithmic_mb_per_call": bytes_per_call / 1e6, "achieved_gbs": gbs,
                               "frac_of_hbm_peak": gbs / peaks["hbm"]}
            mem_ms += bytes_per_call * summ[k]["calls"] / K / (peaks["hbm"] * 1e9) * 1e3
    if "e2e_ctc_fwd_grad" in memory_bound:
        memory_bound["e2e_ctc_fwd_grad"]["note"] = (
            "three launches per head: emission gather (parallel), alpha/beta sweep (T_l sequential frames, two warps "
            "per utterance, 4-8 utterances per CTA), occupancy + gradient (parallel); on its own stream")
    for k in ("e2e_decoder_persist_fwd", "e2e_decoder_persist_bwd"):
        if k in memory_bound:
            memory_bound[k]["note"] = "attention operands only; L2-resident; the kernel also runs the gate GEMMs"

    # ---- dense GEMMs (tensor bound)
    gemm = None
    gemm_ms_at_bound = 0.0
    if shapes:
        tv = summ["e2e_gemm"]
        gemm_ms_at_bound = tv["work"] / K / (peak_tc / split * 1e12) * 1e3
        on_main = {k: v for k, v in shapes.items() if v["main_calls"] > 0} or shapes
        sname, sv = max(on_main.items(), key=lambda kv: kv[1]["main_ms"] if kv[1]["main_calls"] else kv[1]["ms"])
        if sv["main_calls"]:
            sv = {"ms": sv["main_ms"], "calls": sv["main_calls"], "work": sv["main_work"]}
        ach = sv["work"] / (sv["ms"] * 1e-3) / 1e12
        ach_class = tv["work"] / (tv["ms"] * 1e-3) / 1e12
        gemm = {"kernel": "gemm_tc_kernel<%d> %s (the GEMM shape with the largest time on the main stream, incl. its "
                          "operand split passes if any)" % (gemm_mode, sname),
                "bound": "tensor", "achieved": ach, "peak": peak_tc, "unit": "TFLOP/s", "frac": ach / peak_tc,
                "frac_of_peak_over_splits": ach / (peak_tc / split), "mma_products_per_flop": split,
                "issue_ceiling_tflops": peak_tc / units, "frac_of_issue_ceiling": ach / (peak_tc / units),
                "algorithmic_gflop_per_launch": sv["work"] / sv["calls"] / 1e9, "avg_launch_ms": sv["ms"] / sv["calls"],
                "launches_per_step": sv["calls"] / K, "main_stream_ms_per_step": tv["main_ms"] / K,
                "traffic": (traffic.get("gemm_tc_kernel") or {}).get("dram_bytes_per_launch"),
                "class_aggregate": {"kernel": "all e2e_gemm calls (every dense projection incl. dX/dW, %d shapes)"
                                              % len(shapes),
                                    "algorithmic_gflop_per_step": tv["work"] / K / 1e9, "achieved": ach_class,
                                    "frac": ach_class / peak_tc, "frac_of_peak_over_splits": ach_class / (peak_tc / split),
                                    "avg_launch_ms": tv["ms"] / tv["calls"]},
                "top_shapes": [{"shape": k, "ms_per_step": v["ms"] / K, "on_main_stream": v["main_calls"] > 0,
                                "tflops": v["work"] / (v["ms"] * 1e-3) / 1e12}
                               for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:6]],
                "note": "events of GEMMs on the weight-gradient / CTC side streams overlap the main stream and wait "
                        "for SMs held by the persistent kernels: the class total exceeds its share of the step"}

    # ---- the dominant kernel: the kernel with the largest time ON THE MAIN STREAM (a GEMM shape counts as one kernel)
    cand = {k: v["main_ms"] for k, v in summ.items() if k != "e2e_gemm"}
    for k, v in shapes.items():
        cand[k] = v["main_ms"]
    name = max(cand, key=cand.get) if cand else None
    if name in seq and name in flop_per_step:
        v, d = summ[name], seq[name]
        launches = max(v["calls"], 1)
        roof = {"kernel": "%s (entry point %s; the kernel with the largest time on the main stream: %.2f ms = %.0f %% "
                          "of the step in %d launches)" % (d["kernel"], name, v["main_ms"] / K, 100 * d["share_of_step"],
                                                            launches // K),
                "bound": "tensor", "achieved": d["tensor_tflops"], "peak": peak_tc, "unit": "TFLOP/s",
                "frac": d["frac_of_tensor_peak"],
                "traffic": (traffic.get(d["kernel"].split("<")[0]) or {}).get("dram_bytes_per_launch"),
                "traffic_source": (traffic.get(d["kernel"].split("<")[0]) or {}).get("source"),
                "traffic_note": "dram__bytes_read + dram__bytes_write of the FIRST launch ncu captured (one encoder layer: the "
                                "launches of a step differ in T_l; profiles/r2_ncu_full_summary.txt), not of the average launch",
                "peak_source": peaks["source"] + ", sustained bf16 (the kernel sits inside a long step)",
                "algorithmic_gflop_per_launch": flop_per_step[name] * v["work"] / launches / 1e9,
                "avg_launch_ms": v["ms"] / launches, "launches_per_step": launches / K,
                "us_per_timestep": d["us_per_timestep"],
                "latency_floor": {"cycles_per_timestep": d.get("floor_cycles"), "us_per_timestep": d.get("floor_us"),
                                  "frac": d.get("frac_of_latency_floor"), "sm_mhz": sm_mhz,
                                  "source": "scratch/ubench.cu + in-kernel clock64 stamps (DESIGN.md 4.1/4.2)"},
                "note": "a sequential kernel: T_l dependent timesteps of a [B x H] . [H x 4H] product per direction, M = "
                        "16-row tiles on mma.sync; its ALGORITHMIC tensor rate is far below the tensor peak by "
                        "construction, so the fraction that says how good the kernel is is latency_floor.frac = "
                        "(timesteps x floor) / measured"}
    elif gemm is not None and name in shapes:
        roof = dict(gemm)
        roof["peak_source"] = peaks["source"] + ", sustained bf16"
    else:
        roof = {"kernel": name, "bound": "hbm", "achieved": memory_bound.get(name, {}).get("achieved_gbs"),
                "peak": peaks["hbm"], "unit": "GB/s", "frac": memory_bound.get(name, {}).get("frac_of_hbm_peak"),
                "traffic": (traffic.get(name) or {}).get("dram_bytes_per_launch"), "peak_source": peaks["source"]}

    at_bound = {"dense_gemms_ms": gemm_ms_at_bound, "sequential_floors_ms": floor_ms, "memory_bound_ms": mem_ms}
    pct = {"value": sum(at_bound.values()) / ms_per_step, "terms_ms": at_bound, "step_ms": ms_per_step,
           "definition": "SURVEY.md 8(d): sum over kernels of the algorithmic time at the kernel's bound / measured step: "
                         "GEMM FLOPs at peak/%d (mode %s), sequential timesteps x latency floor, attention / CE / CTC "
                         "bytes at the HBM peak" % (split, {0: "fp32", 1: "tf32x3", 2: "bf16", 3: "bf16x2", 4: "f16x2"}[gemm_mode])}
    return {"roofline": roof, "breakdown": breakdown, "sequential_kernels": seq, "memory_bound_kernels": memory_bound,
            "gemm": gemm, "pct_of_roofline": pct}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", default="cfg2", help="cfg2 (default, BASELINE configs[1]) | cfg4 | cfg5 | cfg1")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's batch PER GPU (default); strong: the config's batch split over the GPUs")
    ap.add_argument("--gemm", default="f16x2", help="fp32 | tf32x3 | f16x2 | bf16x2 | bf16 (tf32x3 and f16x2 are the fp32-accurate tensor-core modes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-beam", action="store_true", help="skip the beam-decode utt/s side measurement")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of its CUDA-graph replay")
    ap.add_argument("--dropout", action="store_true",
                    help="the reference's default output dropout: out_prob = out_prob_dec = 0.9 (captured in the graph)")
    ap.add_argument("--defaults", action="store_true",
                    help="the reference's default training settings: --dropout plus samp_prob = 0.1 (scheduled sampling "
                         "realises its ids from the host: eager step)")
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
    if args.scaling == "strong" and world > 1:
        # strong scaling: the config's batch is the GLOBAL batch, utterances dealt round-robin to the ranks
        if cfg.B % world:
            raise SystemExit("bench.py: --scaling strong needs the batch (%d) divisible by the GPUs (%d)" % (cfg.B, world))
        batch = edist.shard_batch(synth.make_batch(cfg), rank, world)
        cfg = synth.get_config(args.config, B=cfg.B // world)
    else:
        batch = synth.make_batch(cfg, seed=synth.DATA_SEED + rank)      # weak scaling: every rank its own batch
    frames = int(batch["logmel_len"].sum())
    reducer = edist.GradAllReducer() if world > 1 else None
    model = build_model(cfg, weights, device=dev, reducer=reducer)
    settings = {}
    if args.dropout or args.defaults:
        # encoder.py:24, decoder.py:25: the keep probabilities the reference trains with
        model.params.encoder_params.out_prob = 0.9
        for d_ in model.params.decoder_params.values():
            d_.out_prob_dec = 0.9
        settings["dropout"] = "out_prob = out_prob_dec = 0.9 (reference defaults)"
    if args.defaults:
        for d_ in model.params.decoder_params.values():
            d_.samp_prob = 0.1                                      # decoder.py:30
        settings["samp_prob"] = 0.1
        args.no_graph = True

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
            model.run_step(prepared=prepared)       # the aborted capture dropped the step's results: redo one
    if use_graph:
        for _ in range(W):
            gs.step()
        ms, host_busy_ms, n_tail, clocks, _ = timed_loop(gs.step, False)
        launches = n_tail + gs.launches_per_step * K         # kernels in the replayed graph + the eager tail
        step_host = lambda: gs.step(batch)
        mode = ("CUDA graph replay of fwd+bwd%s (%d kernels) + eager clip"
                % (" + gradient all-reduce spans" if world > 1 else "", gs.launches_per_step))
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
        # The captured step holds NCCL collectives as graph nodes: destroying the process group (or letting the
        # interpreter tear it down) while such a graph exists blocks forever.  All measurements are done: the other
        # ranks leave right here, rank 0 reports and leaves the same way.
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        if rank != 0:
            os._exit(0)
    if rank != 0:
        return
    # ---------------- per-kernel-class breakdown and roofline (CUDA events of the eager timed pass)
    raw = prof.summary(main_stream=torch.cuda.current_stream().cuda_stream)
    report = roofline_report(raw, cfg, peaks, K, ms / K, clocks, ops.get_gemm_mode())
    roof, breakdown, extra, memory_bound = (report["roofline"], report["breakdown"], report["sequential_kernels"],
                                            report["memory_bound_kernels"])

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": {0: "f32", 1: "tf32x3", 2: "bf16", 3: "bf16x2", 4: "f16x2"}[ops.get_gemm_mode()], "data": "synthetic",
        "config": workload_config(args.config, world, batch_per_gpu=cfg.B, settings=settings),
        "frames_per_step_per_gpu": frames,
        "padded_frames_per_step_per_gpu": cfg.B * cfg.T, "loss": loss_val,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "launches_per_step": launches / K,
        "step_mode": mode, "host_enqueue_ms_per_step": host_busy_ms,
        "eager": {"ms_per_step": ms_eager / K, "host_enqueue_ms_per_step": host_eager_ms,
                  "note": "same K steps launched kernel by kernel; the per-kernel events of breakdown/roofline come "
                          "from this pass (a graph replay runs the identical kernels)"},
        "roofline": roof, "pct_of_roofline": report["pct_of_roofline"], "breakdown": breakdown,
        "sequential_kernels": extra, "memory_bound_kernels": memory_bound, "gemm": report["gemm"],
    }
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample (the whole default run must end within minutes): 8 of the batch's utterances, all host cores,
        # median of 3 after 1 warm-up; `bench.py --impl reference` times the FULL batch (median of 5)
        cores = os.cpu_count() or 1
        sample_B = min(8, cfg.B)
        fps, sec, fr, _ = cpu_reference_step(args.config, sample_B, 3, 1)
        line["cpu_baseline"] = {
            "value": fps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%s shapes, batch %d of %d utterances (%d valid frames), float32 NumPy restatement "
                      "(oracle/model.py; restated reference, not TF), all host cores, median of 3 steps after 1 warm-up, "
                      "%.1f s/step; full batch: bench.py --impl reference" % (args.config, sample_B, cfg.B, fr, sec)}
    if world == 1 and not args.no_beam and args.config == "cfg2":
        line["beam_decode"] = beam_decode_rate(args.config, dev, cpu=not args.no_cpu_baseline)
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        os._exit(0)


# FP64 ceilings of a B200 of this pool, measured with scratch/ubench_f64.cu (profiles/r4_ubench_f64.txt): mma.sync
# m8n8k4.f64 issues every 16.1 cycles per SM sub-partition = 37.0 TFLOP/s at 1.965 GHz; dependent-free DFMA 33.3 TFLOP/s
FP64_DMMA_PEAK_TFLOPS = 37.0
FP64_PEAK_SOURCE = "measured: scratch/ubench_f64.cu, profiles/r4_ubench_f64.txt (DMMA m8n8k4 37.0 TF/s, DFMA 33.3 TF/s)"


def beam_decode_rate(cfg_name, dev, n_utts=256, beam=10, cpu=True):
    """Second half of BASELINE.json's metric: beam-search decoding (BASELINE configs[2]: beam width 10, a synthetic
    eval batch of 256 utterances with T_enc in [50, 88]) in utterances/s through BeamSearch.decode_batch, next to the
    CPU restatement of beam_search.py (serial over utterances like eval_model.py:194-195) timed on a 4-utterance
    sample; the ids of ALL utterances are compared with the oracle's (tests/golden/fullsize_beam.npz, generated once
    by tests/golden/gen_fullsize_golden.py -- the full CPU run takes minutes)."""
    import torch
    from e2e_asr_b200 import synth
    from e2e_asr_b200.beam_search import BeamSearch
    cfg = synth.get_config(cfg_name)
    w = synth.make_weights(cfg)
    encs = synth.make_beam_eval_batch(cfg, n_utts)
    sp = BeamSearch.class_params()
    sp.beam_size = beam
    bs = BeamSearch(w, sp, device=dev)
    bs.decode_batch(encs[:8])
    # first call of this batch shape: launched kernel by kernel; second: captures the decoding step in a CUDA graph
    # (decode_batch keeps the buffers and the graph of a batch signature); from the third on it replays -- the steady
    # state of an evaluation loop over equally shaped batches, which is what `value` reports
    times = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = bs.decode_batch(encs)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    dt = min(times[2:])
    steps = max(len(o) for o in out)
    Hd, Hl, D, E, A, V = cfg.Hd, cfg.Hl, 2 * cfg.H, cfg.E, cfg.A, cfg.V
    flop_row = 2.0 * ((E + Hl) * 4 * Hl + (Hd + D) * E + (E + Hd) * 4 * Hd + Hd * A + (Hd + D) * Hd + Hd * V)
    res = {"value": n_utts / dt, "unit": "utt/s", "beam_size": beam, "n_utts": n_utts,
           "mean_output_len": float(np.mean([len(o) for o in out])), "ms_per_decoding_step": dt * 1e3 / steps,
           "fp64_gemm_tflops": flop_row * n_utts * beam * steps / dt / 1e12,
           "first_call_utt_s": n_utts / times[0], "capture_call_utt_s": n_utts / times[1],
           "note": "random-init weights: hypotheses run to the 120-step limit (worst case); wall clock of decode_batch: "
                   "float64 decoder step on all 2560 hypothesis slots (products on the FP64 tensor cores with the concatenated "
                   "operands taken in place, LM-LSTM input half from a token table, BasicLSTM in the product epilogue, attention "
                   "from tabulated exponentials) + device-side k^2 candidate merge, one CUDA-graph replay "
                   "per decoding step (value = steady state with the step graph cached; first_call_utt_s = kernel by "
                   "kernel, capture_call_utt_s = the call that captures), the best sequence per utterance rebuilt from "
                   "back-pointers on the host at the end"}
    # where a decode goes, kernel by kernel (one more eager decode with an event pair around every C-ABI call), and the
    # FP64 products against the measured FP64 tensor-core ceiling of this pool's B200s
    try:
        from e2e_asr_b200 import _lib
        prof = _lib.Profiler()
        _lib.PROFILER = prof
        try:
            bs.decode_batch(encs, use_graph=False)
            torch.cuda.synchronize()
        finally:
            _lib.PROFILER = None
        summ = prof.summary()
        kern = {k: {"ms_per_decode": v["ms"], "calls": v["calls"], "us_per_call": 1e3 * v["ms"] / max(v["calls"], 1)}
                for k, v in sorted(summ.items(), key=lambda kv: -kv[1]["ms"])}
        gemm_ms = sum(v["ms"] for k, v in summ.items() if k.startswith("e2e_gemm_f64"))
        # executed: the embedding half of the LM-LSTM product comes from the per-model token table
        flop_exec = flop_row - 2.0 * E * 4 * Hl
        R = n_utts * beam
        res["roofline"] = {
            "kernel": "gemm_f64_mma_kernel (e2e_gemm_f64d_cat + e2e_gemm_f64d_lstm: the six float64 products of a decoding "
                      "step on all %d hypothesis rows, BasicLSTM in the epilogue)" % R,
            "bound": "tensor (FP64)", "unit": "TFLOP/s",
            "achieved": flop_exec * R * steps / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None,
            "peak": FP64_DMMA_PEAK_TFLOPS, "peak_source": FP64_PEAK_SOURCE,
            "frac": (flop_exec * R * steps / (gemm_ms * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS) if gemm_ms > 0 else None,
            "executed_gflop_per_step": flop_exec * R / 1e9, "algorithmic_gflop_per_step": flop_row * R / 1e9,
            "share_of_kernel_time": gemm_ms / max(sum(v["ms"] for v in summ.values()), 1e-9),
            "note": "event time of an eager decode; the LSTM pointwise math (5 float64 exp / tanh per unit) runs in the same "
                    "kernels' epilogues and is not counted as FLOPs"}
        res["kernels"] = kern
    except Exception as e:          # noqa: BLE001  (a side measurement must not take the bench line down)
        res["roofline"] = {"error": repr(e)}
    gold = os.path.join(ROOT, "tests", "golden", "fullsize_beam.npz")
    if os.path.exists(gold) and n_utts == 256 and beam == 10 and cfg_name == "cfg2":
        g = np.load(gold)
        off = np.concatenate([[0], np.cumsum(g["lens"])])
        same = [bool(np.array_equal(out[u], g["ids"][off[u]:off[u + 1]])) for u in range(n_utts)]
        res["ids_equal"] = bool(all(same))
        res["ids_equal_utterances"] = "%d of %d (oracle ids: tests/golden/fullsize_beam.npz)" % (sum(same), n_utts)
    if cpu:
        from oracle import beam as ob
        n_cpu = 4
        t0 = time.perf_counter()
        ref = [ob.beam_search(w, e, beam_size=beam) for e in encs[:n_cpu]]
        dt1 = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": n_cpu / dt1, "unit": "utt/s", "cores": 1, "kind": "port",
                               "sample": "%d utterances, Python restatement of beam_search.py (oracle/beam.py), serial "
                                         "like eval_model.py:194-195" % n_cpu,
                               "ids_equal": bool(all(np.array_equal(r, o) for r, o in zip(ref, out)))}
    return res


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Full-size parity goldens for BASELINE.json's configs (run once in the build container, minutes of CPU).

The float64 oracle (oracle/model.py, oracle/beam.py -- pinned to the executed reference by the other generators in this
directory) is run ONCE at the sizes BASELINE.json states and its results are stored in a compact form the GPU tests
(tests/test_gpu_fullsize.py) compare the CUDA path with at 1e-4 / bit-exact ids:

  fullsize_cfg2.npz   cfg-2 (configs[1]) at full size: B=64, T=700, F=120, H=256, L=4, U=120, phone+state CTC
  fullsize_cfg4.npz   cfg-4 (configs[3]): first 8 utterances' worth (B=8) at full T=2000
  fullsize_cfg5.npz   cfg-5 (configs[4]): B=16 at full H=512 / L=5 / T=700 (the kernels cfg-5 really takes)
  fullsize_beam.npz   cfg-3 (configs[2]): oracle beam-search ids, k=10, for ALL 256 utterances of the bench's eval batch

Stored per training config: every task loss, total_loss, the global norm, and for EVERY variable of the clipped
gradient its max-abs, L2 norm, four fixed random-sign projections and a strided sample of <= 4096 entries (the whole
gradient would be 43 MB per config); a strided sample of logits rows and of the top encoder states.

  python tests/golden/gen_fullsize_golden.py [cfg2] [cfg4] [cfg5] [beam]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from e2e_asr_b200 import synth  # noqa: E402

# name -> (config name, batch-size override or None)
TRAIN_CASES = {"cfg2": ("cfg2", None), "cfg4": ("cfg4", 8), "cfg5": ("cfg5", 16)}
SAMPLE = 4096
LOGIT_ROWS = 192
N_PROJ = 4
BEAM_UTTS, BEAM_K = 256, 10


def case_config(tag):
    name, B = TRAIN_CASES[tag]
    return synth.get_config(name, **({} if B is None else {"B": B}))


def sample_index(n, limit=SAMPLE):
    """Strided sample of a flattened variable: every ceil(n/limit)-th entry (all of it when n <= limit)."""
    step = max(1, -(-n // limit))
    return np.arange(0, n, step)


def proj_signs(name, n):
    """Fixed random-sign vectors for the projections of variable `name` (seeded by the name)."""
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31)
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 2, size=(N_PROJ, n)).astype(np.float64) * 2.0 - 1.0


def summarise(tag, out, cfg):
    d = {"total_loss": np.float64(out["total_loss"]), "norm": np.float64(out["norm"]),
         "dense_norm": np.float64(out["dense_norm"])}
    for t, l in out["losses"].items():
        d["loss/" + t] = np.float64(l)
    names = sorted(out["clipped"])
    d["names"] = np.array(names)
    for k in names:
        g = np.asarray(out["clipped"][k], np.float64).ravel()
        d["maxabs/" + k] = np.float64(np.abs(g).max())
        d["l2/" + k] = np.float64(np.sqrt((g * g).sum()))
        d["proj/" + k] = proj_signs(k, g.size) @ g
        d["sample/" + k] = g[sample_index(g.size)].astype(np.float32)
    lg = np.asarray(out["logits"]["char"])
    rows = sample_index(lg.shape[0], LOGIT_ROWS)
    d["logit_rows"] = rows
    d["logits"] = lg[rows].astype(np.float32)
    d["logits_maxabs"] = np.float64(np.abs(lg).max())
    st = np.asarray(out["states"][cfg.L])
    d["states_maxabs"] = np.float64(np.abs(st).max())
    idx = sample_index(st.size, 16384)
    d["states_idx"] = idx
    d["states"] = st.ravel()[idx].astype(np.float32)
    return d


def gen_train(tag):
    from oracle import model as om
    cfg = case_config(tag)
    w = synth.make_weights(cfg)
    batch = synth.make_batch(cfg)
    t0 = time.time()
    out = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, dtype=np.float64)
    print("%s: oracle float64 step %.1f s, total_loss %.9f norm %.6f" % (tag, time.time() - t0, out["total_loss"],
                                                                          out["norm"]), flush=True)
    np.savez(os.path.join(HERE, "fullsize_%s.npz" % tag), **summarise(tag, out, cfg))


def _beam_one(args):
    from oracle import beam as ob
    w, enc = args
    return ob.beam_search(w, enc, beam_size=BEAM_K)


def gen_beam():
    import multiprocessing as mp
    cfg = synth.get_config("cfg2")
    w = synth.make_weights(cfg)
    encs = synth.make_beam_eval_batch(cfg, BEAM_UTTS)
    t0 = time.time()
    with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
        ids = pool.map(_beam_one, [(w, e) for e in encs], chunksize=4)
    print("beam: %d utterances in %.1f s" % (len(ids), time.time() - t0), flush=True)
    lens = np.array([len(i) for i in ids], np.int64)
    np.savez(os.path.join(HERE, "fullsize_beam.npz"), lens=lens, ids=np.concatenate(ids).astype(np.int64),
             enc_lens=np.array([e.shape[0] for e in encs], np.int64), beam_size=np.int64(BEAM_K))


if __name__ == "__main__":
    for tag in (sys.argv[1:] or ["cfg2", "cfg4", "cfg5", "beam"]):
        gen_beam() if tag == "beam" else gen_train(tag)

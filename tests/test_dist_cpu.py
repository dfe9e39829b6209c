"""Data-parallel host logic on CPU (gloo, world_size 2): batch sharding, the flat
gradient all-reduce, and the identity it relies on -- with equal per-rank batch
sizes the global gradient is the mean of the rank gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from e2e_asr_b200 import synth
from e2e_asr_b200 import dist as edist
from oracle import model as om


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = edist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    cfg = synth.get_config("tiny_b", B=6)
    weights = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    shard = edist.shard_batch(batch, rank, world)
    assert len(shard["logmel_len"]) == 3 and shard["logmel"].shape[1] == int(shard["logmel_len"].max())
    res = om.train_step(weights, shard, num_layers={"char": 4}, ctc_tasks=cfg.ctc, max_gradient_norm=1e9)
    names = sorted(res["grads"].keys())
    flat = torch.from_numpy(np.concatenate([res["grads"][n].reshape(-1) for n in names]))
    red = edist.GradAllReducer(bucket_elems=1000)       # several buckets
    red.stream = None                                    # CPU tensors: plain path
    # the overlapped protocol of the training step: spans reduced as "their backward finishes" (out of memory order),
    # the remainder plus one trailing scalar in finish(); the buffer then holds the SUM over ranks
    n = flat.numel()
    buf = torch.cat([flat, torch.tensor([float(rank + 1)], dtype=flat.dtype)])
    red.begin_step(buf)
    red.grad_ready([buf[n - 500:n - 200], buf[n - 200:n]])          # "decoder": two adjacent views, one span
    red.grad_ready([buf[300:n - 700]])                                # an "encoder layer"
    red.grad_ready([buf[n - 600:n - 400]])                            # overlaps a reduced span: must be ignored
    rest = red.finish(n, extra=1)
    assert rest == [(0, 300), (n - 700, n - 500), (n, n + 1)], rest
    assert float(buf[n]) == 3.0                                       # 1 + 2: the scalar rode along
    flat2 = buf[:n] / world
    red.allreduce_mean(flat)
    assert torch.equal(flat, flat2)
    loss = torch.tensor([res["total_loss"]], dtype=torch.float64)
    red.allreduce_sum(loss)
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
        np.save(os.path.join(out_dir, "loss.npy"), loss.numpy() / world)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_mean_equals_global_batch(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cfg = synth.get_config("tiny_b", B=6)
    weights = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(weights, batch, num_layers={"char": 4}, ctc_tasks=cfg.ctc, max_gradient_norm=1e9)
    names = sorted(ref["grads"].keys())
    flat_ref = np.concatenate([ref["grads"][n].reshape(-1) for n in names])
    flat = np.load(os.path.join(str(tmp_path), "flat.npy"))
    np.testing.assert_allclose(flat, flat_ref, rtol=1e-9, atol=1e-12)
    assert abs(float(np.load(os.path.join(str(tmp_path), "loss.npy"))[0]) - ref["total_loss"]) < 1e-10


def test_shard_batch_partitions_every_utterance_once():
    cfg = synth.get_config("tiny_b", B=7)
    batch = synth.make_batch(cfg)
    seen = []
    for r in range(3):
        sh = edist.shard_batch(batch, r, 3)
        seen += list(sh["utt_id"])
        for k in ("logmel", "char", "char_len", "logmel_len"):
            assert len(sh[k]) == len(sh["utt_id"])
    assert sorted(seen) == sorted(batch["utt_id"])

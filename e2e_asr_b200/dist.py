"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch through
torch.distributed (gloo on CPU for the host-logic tests).

The reference has no data parallelism at all; the loss is a
batch mean of per-utterance terms (losses.py:32-35), so with equal per-rank batch
sizes the global gradient is the mean of the rank gradients: ONE allreduce(sum) of
the flat gradient buffer, scaled by 1/n.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        # a mismatched collective must abort the job, not hang it
        import datetime
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("E2E_DIST_TIMEOUT_S", "180"))))
    return rank, world, local


def shard_batch(batch, rank, world):
    """Utterance-level sharding of a global batch: rank r takes rows r::world
    (length-interleaved, so ranks see similar total frames)."""
    if world == 1:
        return batch
    out = {}
    n = len(batch["logmel_len"])
    idx = list(range(rank, n, world))
    for k, v in batch.items():
        out[k] = v[idx]
    # drop padding that no longer belongs to any row of the shard
    t = int(out["logmel_len"].max())
    out["logmel"] = out["logmel"][:, :t]
    return out


class GradAllReducer(object):
    """Averages the flat gradient buffer across ranks.

    `allreduce_mean(flat)` issues one NCCL allreduce per bucket on a side stream and
    makes the compute stream wait for it, so it overlaps whatever the compute stream
    still has queued (the remaining backward when called per bucket)."""

    def __init__(self, group=None, bucket_elems=16 * 1024 * 1024):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bucket_elems = bucket_elems
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None

    def allreduce_sum(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_mean(self, flat):
        if self.world_size == 1:
            return flat
        n = flat.numel()
        if self.stream is None:
            for o in range(0, n, self.bucket_elems):
                chunk = flat[o:o + self.bucket_elems]
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
                chunk.div_(self.world_size)
            return flat
        from ._lib import call
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            for o in range(0, n, self.bucket_elems):
                chunk = flat[o:o + self.bucket_elems]
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
                call("e2e_scale", chunk.numel(), chunk, None, 1.0 / self.world_size)
        cur.wait_stream(self.stream)
        return flat

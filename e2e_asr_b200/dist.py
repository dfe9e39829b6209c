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
        # the gradient spans are a few MB each and run beside the latency-bound recurrences, whose cluster CTAs need whole
        # SMs: a small NCCL grid is enough for them and leaves the SMs alone (measured at N = 2: 4 / 8 / 16 / default
        # CTAs -> 12.67 / 12.45 / 12.51 / 12.56 ms per step)
        os.environ.setdefault("NCCL_MAX_CTAS", "8")
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
    """Sums the flat gradient buffer across ranks, in the order the gradients become final during backward.

    The parameters of the model live in ONE flat buffer laid out in creation order: encoder layer 1 .. L, auxiliary
    heads, decoder (variables.py).  During backward the decoder's gradients are final first, then encoder layer L ..
    1; each group's weight-gradient GEMMs run on a side stream and `grad_ready(views, stream)` (the hook of
    ops.set_grad_ready_hook) starts the all-reduce of THAT span on the reducer's stream as soon as the side stream
    has produced it -- overlapping the remaining backward.  `finish()` reduces whatever was not
    covered (auxiliary heads, plain-autograd paths) together with `extra` trailing floats (scalars that ride along,
    e.g. the IndexedSlices norm term) and joins the compute stream.  The buffer then holds the SUM over ranks; the
    1/n is folded into the clipping kernel (e2e_clip_by_norm pre_scale).  All of it is plain stream-ordered work:
    inside a CUDA-graph capture the collectives become graph nodes, so a replayed step contains its all-reduces.
    """

    def __init__(self, group=None, bucket_elems=16 * 1024 * 1024):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bucket_elems = bucket_elems
        self.stream = torch.cuda.Stream() if torch.cuda.is_available() else None
        self.flat = None
        self.covered = []
        self.collectives = 0

    # ---- overlapped, readiness-ordered path
    def begin_step(self, flat):
        """`flat`: the whole flat gradient buffer (views passed to grad_ready are spans of it)."""
        self.flat = flat
        self.covered = []

    def _reduce_span(self, lo, hi, after=None):
        if hi <= lo:
            return
        chunk = self.flat[lo:hi]
        if self.stream is None:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.stream.wait_stream(after if after is not None else torch.cuda.current_stream())
            with torch.cuda.stream(self.stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        self.collectives += 1

    def grad_ready(self, views, stream=None):
        """The gradient views (spans of the flat buffer) are final once `stream` reaches this point."""
        if self.world_size == 1 or self.flat is None or not views:
            return
        base = self.flat.storage_offset()
        lo = min(v.storage_offset() for v in views) - base
        hi = max(v.storage_offset() + v.numel() for v in views) - base
        # the span between the first and the last view must belong to this group alone (contiguous creation order)
        if any(not (hi <= a or b <= lo) for a, b in self.covered):
            return                                  # overlaps something already reduced: leave it to finish()
        self.covered.append((lo, hi))
        self._reduce_span(lo, hi, after=stream)

    def finish(self, used, extra=0):
        """Reduces every span of [0, used + extra) not covered by grad_ready and makes the current stream wait for all
        collectives of the step.  Returns the list of spans reduced here."""
        if self.world_size == 1 or self.flat is None:
            return []
        rest, pos = [], 0
        for a, b in sorted(self.covered):
            if a > pos:
                rest.append((pos, a))
            pos = max(pos, b)
        if pos < used + extra:
            rest.append((pos, used + extra))
        for a, b in rest:
            for o in range(a, b, self.bucket_elems):
                self._reduce_span(o, min(b, o + self.bucket_elems))
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.covered = []
        return rest

    # ---- plain paths
    def allreduce_sum(self, t):
        if self.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_mean(self, flat):
        """One blocking pass: sum over ranks, divided by n (host-logic tests and non-overlapped callers)."""
        if self.world_size == 1:
            return flat
        self.begin_step(flat)
        self.finish(flat.numel())
        flat.div_(self.world_size)
        return flat

"""Length-bucketed padded batching: the host-side data path in front of
Seq2SeqModel.get_batch, with the reference's batch layout and bucket scheme but without TFRecords.

Reference: speech_dataset.py:43-60 yields dicts {"logmel" [B,T,F] f32, "logmel_len" [B] i64, "char" [B,U+1] i64 (GO ..
EOS, zero = PAD padded), "char_len" [B] i64 (= number of targets, excludes GO), "phone", "phone_len", "utt_id"} through
`padded_batch` (pad to the longest of the batch); train.py:44,108-119 keeps one dataset per pre-computed LENGTH BUCKET
with its own batch size (`buck_batch_size = [128, 128, 64, 64, 32]`: longer utterances, smaller batches) and shuffles
within a bucket.  Here the buckets are computed from the frame counts instead of being read from file names.

With `pad_to_bucket=True` every batch of a bucket is padded to the bucket's upper frame bound, so a bucket has ONE
logmel shape -- the unit a captured step (GraphedStep) is built for."""
import numpy as np

from .data_utils import PAD_ID


class BucketedBatcher(object):
    """boundaries: ascending upper frame bounds of the buckets (the last bucket takes everything longer unless
    `drop_longer`); batch_sizes: one per bucket (len(boundaries) + 1 without drop_longer)."""

    def __init__(self, boundaries, batch_sizes, pad_to_bucket=False, drop_longer=False, seed=0,
                 label_keys=("char", "phone")):
        self.boundaries = [int(b) for b in boundaries]
        assert self.boundaries == sorted(self.boundaries) and len(set(self.boundaries)) == len(self.boundaries)
        n_buckets = len(self.boundaries) + (0 if drop_longer else 1)
        assert len(batch_sizes) == n_buckets, "one batch size per bucket (%d)" % n_buckets
        self.batch_sizes = [int(b) for b in batch_sizes]
        self.pad_to_bucket, self.drop_longer = pad_to_bucket, drop_longer
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.label_keys = tuple(label_keys)

    def bucket_of(self, n_frames):
        """Index of the bucket holding an utterance of n_frames frames, or None if it is dropped."""
        b = int(np.searchsorted(self.boundaries, n_frames, side="left"))
        if b == len(self.boundaries) and self.drop_longer:
            return None
        return b

    def collate(self, utts, pad_frames=None):
        """padded_batch of speech_dataset.py:53-57 over a list of utterance dicts (logmel [T,F], <label> [n] ids with
        GO first and EOS last, utt_id)."""
        B = len(utts)
        F = utts[0]["logmel"].shape[1]
        lens = np.array([u["logmel"].shape[0] for u in utts], np.int64)
        T = int(lens.max()) if pad_frames is None else int(pad_frames)
        assert T >= lens.max()
        logmel = np.zeros((B, T, F), np.float32)
        for i, u in enumerate(utts):
            logmel[i, :lens[i]] = u["logmel"]
        batch = {"logmel": logmel, "logmel_len": lens,
                 "utt_id": np.array([u.get("utt_id", "utt%d" % i) for i, u in enumerate(utts)])}
        for key in self.label_keys:
            if key not in utts[0]:
                continue
            n = np.array([len(u[key]) for u in utts], np.int64)
            ids = np.full((B, int(n.max())), PAD_ID, np.int64)
            for i, u in enumerate(utts):
                ids[i, :n[i]] = u[key]
            batch[key] = ids
            batch[key + "_len"] = n - 1            # number of targets: the ids without GO (cint_len of the records)
        return batch

    def batches(self, utterances, shuffle=True, drop_remainder=False):
        """Yields (bucket index, batch dict).  Utterances are shuffled within their bucket (train.py:113) and the
        buckets' batches are interleaved in random order when `shuffle`."""
        buckets = [[] for _ in self.batch_sizes]
        for u in utterances:
            b = self.bucket_of(u["logmel"].shape[0])
            if b is not None:
                buckets[b].append(u)
        plan = []
        for b, items in enumerate(buckets):
            order = self.rng.permutation(len(items)) if shuffle else np.arange(len(items))
            bs = self.batch_sizes[b]
            for i in range(0, len(items), bs):
                chunk = [items[j] for j in order[i:i + bs]]
                if len(chunk) == bs or not drop_remainder:
                    plan.append((b, chunk))
        if shuffle:
            plan = [plan[i] for i in self.rng.permutation(len(plan))]
        for b, chunk in plan:
            pad = self.boundaries[b] if self.pad_to_bucket and b < len(self.boundaries) else None
            yield b, self.collate(chunk, pad_frames=pad)

"""Length-bucketed padded batching: the reference's batch layout (speech_dataset.py:43-60) and
bucket scheme (train.py:44,108-119)."""
import numpy as np

from e2e_asr_b200.batching import BucketedBatcher
from e2e_asr_b200.data_utils import EOS_ID, GO_ID, PAD_ID


def make_utts(n, seed=0, F=5):
    rng = np.random.default_rng(seed)
    utts = []
    for i in range(n):
        T = int(rng.integers(3, 120))
        U = int(rng.integers(1, 9))
        ids = np.concatenate([[GO_ID], rng.integers(3, 30, size=U - 1), [EOS_ID]]).astype(np.int64)
        utts.append({"logmel": rng.standard_normal((T, F)).astype(np.float32), "char": ids, "utt_id": "u%03d" % i})
    return utts


def test_every_utterance_lands_once_in_the_right_bucket_with_the_reference_layout():
    utts = make_utts(203)
    bb = BucketedBatcher([30, 60, 90], [16, 8, 4, 2], seed=1)
    seen = {}
    for b, batch in bb.batches(utts):
        B, T, F = batch["logmel"].shape
        assert B <= bb.batch_sizes[b] and batch["logmel"].dtype == np.float32
        assert batch["logmel_len"].dtype == np.int64 and batch["char"].dtype == np.int64
        assert T == batch["logmel_len"].max()                       # padded_batch: pad to the longest of the batch
        lo = bb.boundaries[b - 1] if b > 0 else 0
        assert all(lo < n for n in batch["logmel_len"]) or b == 0
        if b < len(bb.boundaries):
            assert batch["logmel_len"].max() <= bb.boundaries[b]
        for i in range(B):
            u = next(x for x in utts if x["utt_id"] == batch["utt_id"][i])
            n, m = batch["logmel_len"][i], batch["char_len"][i]
            assert np.array_equal(batch["logmel"][i, :n], u["logmel"]) and not batch["logmel"][i, n:].any()
            assert m == len(u["char"]) - 1                          # targets exclude GO
            assert np.array_equal(batch["char"][i, :m + 1], u["char"]) and (batch["char"][i, m + 1:] == PAD_ID).all()
            assert batch["char"][i, 0] == GO_ID and batch["char"][i, m] == EOS_ID
            seen[u["utt_id"]] = seen.get(u["utt_id"], 0) + 1
    assert len(seen) == len(utts) and set(seen.values()) == {1}


def test_pad_to_bucket_gives_one_shape_per_bucket_and_drop_options():
    utts = make_utts(150, seed=3)
    bb = BucketedBatcher([40, 80], [8, 4], pad_to_bucket=True, drop_longer=True, seed=2)
    shapes = {}
    n = 0
    for b, batch in bb.batches(utts, drop_remainder=True):
        shapes.setdefault(b, set()).add(batch["logmel"].shape)
        n += batch["logmel"].shape[0]
    assert all(len(s) == 1 for s in shapes.values())
    assert shapes[0] == {(8, 40, 5)} and shapes[1] == {(4, 80, 5)}
    kept = [u for u in utts if u["logmel"].shape[0] <= 80]
    assert n <= len(kept) and n >= len(kept) - (8 - 1) - (4 - 1)
    assert bb.bucket_of(81) is None and bb.bucket_of(40) == 0 and bb.bucket_of(41) == 1
    # deterministic for a seed, different order for another
    a = [tuple(batch["utt_id"]) for _, batch in BucketedBatcher([40, 80], [8, 4, 2], seed=5).batches(utts)]
    b = [tuple(batch["utt_id"]) for _, batch in BucketedBatcher([40, 80], [8, 4, 2], seed=5).batches(utts)]
    c = [tuple(batch["utt_id"]) for _, batch in BucketedBatcher([40, 80], [8, 4, 2], seed=6).batches(utts)]
    assert a == b and a != c

"""SpeechDataset (reference speech_dataset.py:1-60) without TensorFlow: TFRecord framing, SequenceExample parsing with
the reference's schema, shuffle / padded_batch / end-of-data semantics.  No TensorFlow-written file exists in this image
(parity unpinned against one): the parser is checked against bytes assembled BY HAND from the public wire format
(record_writer.h, example.proto, feature.proto), independently of the module's own writer."""
import struct

import numpy as np
import pytest

from e2e_asr_b200 import speech_dataset as sd
from e2e_asr_b200.base_params import Bunch
from e2e_asr_b200.tf_checkpoint import crc32c, mask_crc


def ld(tag, payload):
    """Length-delimited protobuf field with a one-byte tag and a one-byte length."""
    assert len(payload) < 128
    return bytes([tag, len(payload)]) + payload


def hand_assembled_example():
    # Feature{int64_list(3){value(1) packed}} : 0x1a len { 0x0a len varints }
    i64 = lambda *v: ld(0x1a, ld(0x0a, bytes(v)))
    byt = lambda b: ld(0x0a, ld(0x0a, b))
    f_packed = lambda *v: ld(0x12, ld(0x0a, struct.pack("<%df" % len(v), *v)))
    f_unpacked = lambda *v: ld(0x12, b"".join(bytes([0x0d]) + struct.pack("<f", x) for x in v))   # wire type 5
    entry = lambda k, val: ld(0x0a, ld(0x0a, k) + ld(0x12, val))                                  # map entry {1: key, 2: value}
    context = entry(b"segment", byt(b"sw02001-A_000098")) + entry(b"logmel_len", i64(2)) + \
        entry(b"cint_len", i64(2)) + entry(b"pint_len", i64(1))
    flist = lambda *feats: b"".join(ld(0x0a, f) for f in feats)                                    # FeatureList{1: Feature}
    lists = entry(b"logmel", flist(f_packed(0.5, -1.25), f_unpacked(3.0, 4.5))) + \
        entry(b"cint", flist(i64(1), i64(0x85, 0x01), i64(2))) + entry(b"pint", flist(i64(1), i64(2)))   # 0x85 0x01 = 133
    return ld(0x0a, context) + bytes([0x12, 0x80 | (len(lists) & 0x7F), len(lists) >> 7]) + lists         # 2-byte length


def test_parses_hand_assembled_sequence_example(tmp_path):
    proto = hand_assembled_example()
    ds = sd.SpeechDataset(Bunch(batch_size=2, feat_length=2), [], isTraining=False)
    inst = ds.get_instance(proto)
    np.testing.assert_array_equal(inst["logmel"], np.array([[0.5, -1.25], [3.0, 4.5]], np.float32))
    np.testing.assert_array_equal(inst["char"], [1, 133, 2])
    np.testing.assert_array_equal(inst["phone"], [1, 2])
    assert (int(inst["logmel_len"]), int(inst["char_len"]), int(inst["phone_len"])) == (2, 2, 1)
    assert inst["utt_id"] == b"sw02001-A_000098"
    # TFRecord framing by hand: length, masked crc of the length bytes, data, masked crc of the data
    head = struct.pack("<Q", len(proto))
    rec = head + struct.pack("<I", mask_crc(crc32c(head))) + proto + struct.pack("<I", mask_crc(crc32c(proto)))
    path = tmp_path / "one.tfrecord"
    path.write_bytes(rec + rec)
    assert [r for r in sd.read_records(str(path))] == [proto, proto]
    bad = bytearray(rec)
    bad[20] ^= 0x40
    path.write_bytes(bytes(bad))
    with pytest.raises(ValueError, match="corrupted"):
        list(sd.read_records(str(path)))
    path.write_bytes(rec[:-3])
    with pytest.raises(ValueError, match="truncated"):
        list(sd.read_records(str(path)))
    # the module's own writer emits the same bytes for the same utterance (packed encoding everywhere)
    again = sd.parse_sequence_example(sd.make_sequence_example(
        {"utt_id": "sw02001-A_000098", "logmel": inst["logmel"], "char": inst["char"], "phone": inst["phone"]}))
    assert again[0]["segment"][1] == [b"sw02001-A_000098"] and int(again[0]["cint_len"][1][0]) == 2


def _utterances(n, F, rng):
    utts = []
    for i in range(n):
        T, nc, npn = int(rng.integers(3, 40)), int(rng.integers(2, 9)), int(rng.integers(2, 6))
        utts.append({"utt_id": "utt%03d" % i, "logmel": rng.standard_normal((T, F)).astype(np.float32),
                     "char": np.r_[1, rng.integers(3, 50, nc - 2), 2], "phone": np.r_[1, rng.integers(3, 20, npn - 2), 2]})
    return utts


def test_round_trip_padded_batches_and_end_of_data(tmp_path):
    rng = np.random.default_rng(0)
    F = 5
    utts = _utterances(11, F, rng)
    files = [str(tmp_path / "train_1k.0.a"), str(tmp_path / "train_1k.0.b")]
    sd.write_tfrecord(files[0], utts[:6])
    sd.write_tfrecord(files[1], utts[6:])
    ds = sd.SpeechDataset(Bunch(batch_size=4, feat_length=F), files, isTraining=False)
    assert ds.data_set is ds
    it = ds.data_iter
    for _pass in range(2):                                   # initializer restarts the pass (train.py:370)
        it.initializer()
        seen = []
        for start in (0, 4, 8):
            b = it.get_next()
            chunk = utts[start:start + 4]
            assert [u.decode() for u in b["utt_id"]] == [u["utt_id"] for u in chunk]
            T = max(u["logmel"].shape[0] for u in chunk)
            assert b["logmel"].shape == (len(chunk), T, F) and b["logmel"].dtype == np.float32
            for i, u in enumerate(chunk):
                n = u["logmel"].shape[0]
                np.testing.assert_array_equal(b["logmel"][i, :n], u["logmel"])
                assert not b["logmel"][i, n:].any()
                np.testing.assert_array_equal(b["char"][i, :len(u["char"])], u["char"])
                assert not b["char"][i, len(u["char"]):].any()                 # PAD_ID = 0
                assert b["char_len"][i] == len(u["char"]) - 1 and b["phone_len"][i] == len(u["phone"]) - 1
                assert b["logmel_len"][i] == n
            seen += list(b["utt_id"])
        with pytest.raises(sd.OutOfRangeError):
            it.get_next()
        assert len(seen) == 11
    with pytest.raises(ValueError, match="feat_length"):
        sd.SpeechDataset(Bunch(batch_size=4, feat_length=F + 1), files, isTraining=False).data_iter.get_next()


def test_training_iterator_shuffles_but_keeps_every_utterance(tmp_path):
    rng = np.random.default_rng(1)
    utts = _utterances(40, 3, rng)
    path = str(tmp_path / "train_1k.1.a")
    sd.write_tfrecord(path, utts)
    ds = sd.SpeechDataset(Bunch(batch_size=8, feat_length=3), path, isTraining=True, seed=3)
    ds.SHUFFLE_BUFFER = 16                                   # smaller than the data: exercises the refill path
    ids = []
    while True:
        try:
            ids += [u.decode() for u in ds.data_iter.get_next()["utt_id"]]
        except sd.OutOfRangeError:
            break
    assert sorted(ids) == [u["utt_id"] for u in utts] and ids != [u["utt_id"] for u in utts]

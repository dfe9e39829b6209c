"""The hand-written protobuf readers / writers (speech_dataset.py: tf.train.SequenceExample; tf_checkpoint.py:
BundleHeaderProto / BundleEntryProto) against GOOGLE'S protobuf runtime declared with the public TensorFlow schemas
(tests/pb_schema.py): messages encoded by the runtime must parse here, and bytes written here must decode to the same
messages there.  This pins the wire-format handling against an independent implementation; what stays unpinned is a
file written by TensorFlow itself (none exists in this image)."""
import numpy as np
import pytest

pb = pytest.importorskip("google.protobuf")

import pb_schema  # noqa: E402  (tests/ is on sys.path under pytest's rootdir conftest)
from e2e_asr_b200 import speech_dataset as sd  # noqa: E402
from e2e_asr_b200 import tf_checkpoint as tc  # noqa: E402
from e2e_asr_b200.base_params import Bunch  # noqa: E402

M = pb_schema.build()


def _runtime_example(utt_id, logmel, char, phone, lens):
    """tf.train.SequenceExample of the reference's schema (speech_dataset.py:13-33), built by the protobuf runtime."""
    se = M["SequenceExample"]()
    se.context.feature["segment"].bytes_list.value.append(utt_id)
    for name, v in zip(("logmel_len", "cint_len", "pint_len"), lens):
        se.context.feature[name].int64_list.value.append(int(v))
    for row in logmel:
        se.feature_lists.feature_list["logmel"].feature.add().float_list.value.extend(float(x) for x in row)
    for name, ids in (("cint", char), ("pint", phone)):
        fl = se.feature_lists.feature_list[name]
        for i in ids:
            fl.feature.add().int64_list.value.append(int(i))
    return se


@pytest.mark.parametrize("seed", range(4))
def test_runtime_encoded_sequence_example_parses(seed):
    rng = np.random.default_rng(seed)
    T, F = int(rng.integers(1, 300)), int(rng.integers(1, 90))
    logmel = rng.standard_normal((T, F)).astype(np.float32)
    # ids that need 1-, 2-, 5- and 10-byte varints (a negative int64 is ten bytes on the wire)
    char = np.r_[1, rng.integers(3, 100000, int(rng.integers(1, 200))), 2 ** 40 + 7, -3, 2].astype(np.int64)
    phone = np.r_[1, rng.integers(3, 50, int(rng.integers(0, 60))), 2].astype(np.int64)
    se = _runtime_example(b"sw0%d-A" % seed, logmel, char, phone, (T, len(char) - 1, len(phone) - 1))
    ds = sd.SpeechDataset(Bunch(batch_size=2, feat_length=F), [], isTraining=False)
    inst = ds.get_instance(se.SerializeToString())
    np.testing.assert_array_equal(inst["logmel"], logmel)
    np.testing.assert_array_equal(inst["char"], char)
    np.testing.assert_array_equal(inst["phone"], phone)
    assert (int(inst["logmel_len"]), int(inst["char_len"]), int(inst["phone_len"])) == (T, len(char) - 1, len(phone) - 1)
    assert inst["utt_id"] == b"sw0%d-A" % seed
    # deterministic=True sorts the map entries: another legal byte order of the same message
    inst2 = ds.get_instance(se.SerializeToString(deterministic=True))
    np.testing.assert_array_equal(inst2["logmel"], logmel)
    np.testing.assert_array_equal(inst2["char"], char)


def test_writer_output_decodes_to_the_same_message_in_the_runtime():
    rng = np.random.default_rng(7)
    utt = {"utt_id": "sw04099-B_012345", "logmel": rng.standard_normal((37, 80)).astype(np.float32),
           "char": np.r_[1, rng.integers(3, 1000, 140), 2], "phone": np.r_[1, rng.integers(3, 48, 60), 2]}
    got = M["SequenceExample"]()
    got.ParseFromString(sd.make_sequence_example(utt))
    want = _runtime_example(b"sw04099-B_012345", utt["logmel"], utt["char"], utt["phone"],
                            (37, len(utt["char"]) - 1, len(utt["phone"]) - 1))
    assert got == want
    # and the runtime's re-encoding of what we wrote parses back here
    ctx, lists = sd.parse_sequence_example(got.SerializeToString())
    assert ctx["segment"][1] == [b"sw04099-B_012345"] and len(lists["logmel"]) == 37


@pytest.mark.parametrize("shape,dtype_id,shard,offset,size,crc", [
    ((), 1, 0, 0, 4, 0), ((1024, 2048), 1, 0, 1 << 33, 1024 * 2048 * 4, 0xDEADBEEF), ((7,), 9, 3, 12345, 56, 1),
    ((0, 5), 1, 0, 77, 0, 0x80000000)])
def test_bundle_entry_matches_the_runtime(shape, dtype_id, shard, offset, size, crc):
    """BundleEntryProto (tensor_bundle.proto: dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5, crc32c = 6
    fixed32) written here == the runtime's message, and the runtime's bytes parse here (proto3 omits zero fields)."""
    mine = tc._encode_entry(dtype_id, shape, shard, offset, size, crc)
    got = M["BundleEntryProto"]()
    got.ParseFromString(mine)
    want = M["BundleEntryProto"](dtype=dtype_id, shard_id=shard, offset=offset, size=size, crc32c=crc)
    for d in shape:
        want.shape.dim.add().size = d
    if not shape:
        want.shape.SetInParent()
    assert got == want
    fields = tc._parse_proto(want.SerializeToString())
    first = lambda k, default=0: fields[k][0] if k in fields else default
    assert first(1) == dtype_id and first(3) == shard and first(4) == offset and first(5) == size
    assert first(6) == crc
    dims = [tc._parse_proto(d).get(1, [0])[0] for d in tc._parse_proto(fields[2][0]).get(2, [])] if 2 in fields else []
    assert tuple(dims) == tuple(shape)


def test_checkpoint_written_here_is_readable_entry_by_entry_by_the_runtime(tmp_path):
    """Every value of the index table of a checkpoint written by write_checkpoint decodes in the runtime: the header
    (key "") as BundleHeaderProto{num_shards = 1, little endian, version.producer = 1}, the rest as BundleEntryProto
    whose offset / size / shape describe the tensor's bytes in the data shard."""
    rng = np.random.default_rng(3)
    tensors = {"model/encoder/w": rng.standard_normal((5, 7)).astype(np.float32),
               "model/global_step": np.asarray(1234, np.int64),
               "model/rnn_decoder_char/decoder/embedding": rng.standard_normal((11, 3)).astype(np.float32)}
    prefix = str(tmp_path / "asr.ckpt-1234")
    tc.write_checkpoint(prefix, tensors)
    table = dict(tc.read_table(prefix + ".index"))
    hdr = M["BundleHeaderProto"]()
    hdr.ParseFromString(table[b""])
    assert hdr.num_shards == 1 and hdr.endianness == 0 and hdr.version.producer == 1
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    for name, t in tensors.items():
        e = M["BundleEntryProto"]()
        e.ParseFromString(table[name.encode()])
        assert tuple(d.size for d in e.shape.dim) == t.shape and e.size == t.nbytes and e.shard_id == 0
        assert e.dtype == {np.dtype(np.float32): 1, np.dtype(np.int64): 9}[t.dtype]
        raw = data[e.offset:e.offset + e.size]
        assert raw == t.tobytes() and tc.unmask_crc(e.crc32c) == tc.crc32c(raw)
    back = tc.read_checkpoint(prefix)
    for name, t in tensors.items():
        np.testing.assert_array_equal(back[name], t)

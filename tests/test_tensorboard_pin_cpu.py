"""TFRecord framing, CRC32C + masking and the DataType enum pinned against TensorBoard's own implementations (the
TensorFlow team's pure-Python record writer / reader and vendored protos; TensorBoard is in the image, TensorFlow is
not): records written by TensorBoard's RecordWriter must read here, records written here must read in TensorBoard's
PyRecordReader, CRCs must agree on arbitrary bytes, and tf_checkpoint's dtype ids must be types.proto's."""
import numpy as np
import pytest

pytest.importorskip("tensorboard")

from tensorboard.compat.proto import tensor_shape_pb2, types_pb2, versions_pb2  # noqa: E402
from tensorboard.compat.tensorflow_stub import pywrap_tensorflow as tbw  # noqa: E402
from tensorboard.summary.writer.record_writer import RecordWriter  # noqa: E402

from e2e_asr_b200 import speech_dataset as sd  # noqa: E402
from e2e_asr_b200 import tf_checkpoint as tc  # noqa: E402


def _payloads(rng):
    sizes = [0, 1, 7, 127, 128, 4096, 70001]
    return [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in sizes]


def test_crc32c_and_mask_agree_with_tensorboard():
    rng = np.random.default_rng(0)
    for data in _payloads(rng) + [b"123456789"]:
        assert tc.crc32c(data) == tbw.crc32c(data)
        assert tc.mask_crc(tc.crc32c(data)) == tbw.masked_crc32c(data)
        assert tc.unmask_crc(tbw.masked_crc32c(data)) == tbw.crc32c(data)
    assert tc.crc32c(b"123456789") == 0xE3069283                     # the CRC-32C check value


def test_records_written_by_tensorboard_read_here(tmp_path):
    rng = np.random.default_rng(1)
    payloads = _payloads(rng)
    path = tmp_path / "tb.tfrecord"
    with open(path, "wb") as f:
        w = RecordWriter(f)
        for p in payloads:
            w.write(p)
        w.flush()
    assert list(sd.read_records(str(path))) == payloads


def test_records_written_here_read_in_tensorboard(tmp_path):
    rng = np.random.default_rng(2)
    payloads = _payloads(rng)
    path = tmp_path / "mine.tfrecord"
    sd.write_records(str(path), payloads)
    r = tbw.PyRecordReader_New(str(path))
    got = []
    while True:
        try:
            r.GetNext()
        except tbw.errors.OutOfRangeError:
            break
        got.append(r.record())
    assert got == payloads


def test_dtype_ids_are_those_of_types_proto():
    names = {np.float32: "DT_FLOAT", np.float64: "DT_DOUBLE", np.int32: "DT_INT32", np.uint8: "DT_UINT8",
             np.int16: "DT_INT16", np.int8: "DT_INT8", np.int64: "DT_INT64", np.bool_: "DT_BOOL", np.uint16: "DT_UINT16",
             np.float16: "DT_HALF", np.uint32: "DT_UINT32", np.uint64: "DT_UINT64"}
    assert len(tc._DTYPES) == len(names)
    for enum_id, np_type in tc._DTYPES.items():
        assert types_pb2.DataType.Value(names[np_type]) == enum_id


def test_shape_and_version_submessages_decode_in_the_official_protos(tmp_path):
    """The TensorShapeProto inside a BundleEntryProto and the VersionDef inside the BundleHeaderProto written here,
    decoded by the TF-generated message classes TensorBoard vendors."""
    prefix = str(tmp_path / "m.ckpt")
    tc.write_checkpoint(prefix, {"a/b": np.zeros((3, 0, 5), np.float32), "c": np.asarray(2.5, np.float64)})
    table = dict(tc.read_table(prefix + ".index"))
    entry = tc._parse_proto(table[b"a/b"])
    shape = tensor_shape_pb2.TensorShapeProto()
    shape.ParseFromString(entry[2][0])
    assert [d.size for d in shape.dim] == [3, 0, 5] and not shape.unknown_rank
    scalar = tc._parse_proto(table[b"c"])
    shape = tensor_shape_pb2.TensorShapeProto()
    shape.ParseFromString(scalar.get(2, [b""])[0])
    assert len(shape.dim) == 0 and scalar[1][0] == types_pb2.DT_DOUBLE
    header = tc._parse_proto(table[b""])
    ver = versions_pb2.VersionDef()
    ver.ParseFromString(header[3][0])
    assert ver.producer == 1 and ver.min_consumer == 0

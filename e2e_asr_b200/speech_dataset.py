"""SpeechDataset: the reference's input pipeline (speech_dataset.py:1-60) without TensorFlow.

Reads TFRecord files of serialized `tf.train.SequenceExample` protos with the reference's schema

    context        segment (bytes), logmel_len, cint_len, pint_len (int64)                 speech_dataset.py:15-20
    feature_lists  logmel: T x float_list[feat_length]; cint, pint: n x int64_list[1]      speech_dataset.py:21-25

maps every record to {"logmel", "char", "phone", "logmel_len", "char_len", "phone_len", "utt_id"} (:43-45), shuffles
with a 4000-record buffer when training (:51-52) and zero-pads batches of `params.batch_size` records to the longest of
the batch (`padded_batch`, :53-57).  `data_iter.get_next()` returns the batch dict `Seq2SeqModel.get_batch` consumes and
raises `OutOfRangeError` at the end of the data (the reference's end-of-epoch signal, train.py:379);
`data_iter.initialize()` restarts it (`sess.run(data_iter.initializer)`).

File formats (public specifications, restated here; PARITY UNPINNED against files written by TensorFlow, none exist in
this image -- checked against hand-assembled bytes, round trips and corruption detection in tests/):
  * TFRecord framing: u64 length | u32 masked crc32c(length) | data | u32 masked crc32c(data), little endian
    (tensorflow/core/lib/io/record_writer.h); the masked CRC32C is the one of tf_checkpoint.py;
  * protobuf wire format of tensorflow/core/example/example.proto / feature.proto: SequenceExample{1: context Features,
    2: FeatureLists}, Features{1: map<string, Feature>}, Feature{1: BytesList, 2: FloatList, 3: Int64List},
    FloatList / Int64List{1: repeated value, packed or not}, FeatureLists{1: map<string, FeatureList>},
    FeatureList{1: repeated Feature}.
A writer (`write_tfrecord`) produces the same format for synthetic data and the tests.
"""
import struct

import numpy as np

from .data_utils import PAD_ID
from .tf_checkpoint import _field, _get_varint, _parse_proto, _put_varint, crc32c, mask_crc


class OutOfRangeError(Exception):
    """End of the data: what tf.errors.OutOfRangeError signals to the reference's training loop (train.py:379)."""


# ----------------------------------------------------------------------------- TFRecord framing
def read_records(path, verify=True):
    """Yields the payload of every record of a TFRecord file."""
    with open(path, "rb") as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError("%s: truncated record header" % path)
            (n,), (crc_len,) = struct.unpack("<Q", head[:8]), struct.unpack("<I", head[8:])
            if verify and mask_crc(crc32c(head[:8])) != crc_len:
                raise ValueError("%s: corrupted record length" % path)
            data = f.read(n)
            tail = f.read(4)
            if len(data) < n or len(tail) < 4:
                raise ValueError("%s: truncated record" % path)
            if verify and mask_crc(crc32c(data)) != struct.unpack("<I", tail)[0]:
                raise ValueError("%s: corrupted record data" % path)
            yield data


def write_records(path, payloads):
    with open(path, "wb") as f:
        for data in payloads:
            head = struct.pack("<Q", len(data))
            f.write(head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data))))


# ----------------------------------------------------------------------------- Feature protos
def _feature_values(buf):
    """One Feature message -> ("bytes", [bytes]) | ("float", float32 array) | ("int64", int64 array)."""
    msg = _parse_proto(buf)
    if 1 in msg:
        return "bytes", _parse_proto(msg[1][0]).get(1, [])
    if 2 in msg:
        vals = []
        for v in _parse_proto(msg[2][0]).get(1, []):
            # packed: one length-delimited run of little-endian floats; unpacked: one fixed32 per value
            vals.append(np.frombuffer(v, "<f4") if isinstance(v, bytes) else
                        np.frombuffer(struct.pack("<I", v), "<f4"))
        return "float", np.concatenate(vals) if vals else np.zeros(0, np.float32)
    if 3 in msg:
        vals = []
        for v in _parse_proto(msg[3][0]).get(1, []):
            if isinstance(v, bytes):                     # packed varints
                pos = 0
                while pos < len(v):
                    x, pos = _get_varint(v, pos)
                    vals.append(x)
            else:
                vals.append(v)
        arr = np.array(vals, np.uint64).astype(np.int64)  # two's complement for negatives
        return "int64", arr
    return "none", None


def _map_entries(buf):
    """map<string, X> field 1 of a Features / FeatureLists message -> {key: serialized X}."""
    out = {}
    for entry in _parse_proto(buf).get(1, []):
        e = _parse_proto(entry)
        out[e[1][0].decode("utf-8")] = e[2][0] if 2 in e else b""
    return out


def parse_sequence_example(proto):
    """(context {name: (kind, values)}, feature_lists {name: [(kind, values) per step]}) of a SequenceExample."""
    msg = _parse_proto(proto)
    context = {k: _feature_values(v) for k, v in _map_entries(msg[1][0]).items()} if 1 in msg else {}
    lists = {}
    if 2 in msg:
        for k, v in _map_entries(msg[2][0]).items():
            lists[k] = [_feature_values(f) for f in _parse_proto(v).get(1, [])]
    return context, lists


def _enc_feature(kind, values):
    if kind == "bytes":
        body = b"".join(_field(1, 2, _put_varint(len(v)) + v) for v in values)
        return _field(1, 2, _put_varint(len(body)) + body)
    if kind == "float":
        raw = np.asarray(values, "<f4").tobytes()
        body = _field(1, 2, _put_varint(len(raw)) + raw)
        return _field(2, 2, _put_varint(len(body)) + body)
    raw = b"".join(_put_varint(int(v) & 0xFFFFFFFFFFFFFFFF) for v in values)
    body = _field(1, 2, _put_varint(len(raw)) + raw)
    return _field(3, 2, _put_varint(len(body)) + body)


def _enc_map(entries):
    out = b""
    for k, v in entries:
        kb = k.encode("utf-8")
        e = _field(1, 2, _put_varint(len(kb)) + kb) + _field(2, 2, _put_varint(len(v)) + v)
        out += _field(1, 2, _put_varint(len(e)) + e)
    return out


def make_sequence_example(utt):
    """Serialized SequenceExample of one utterance dict {"utt_id", "logmel" [T,F], "char" [n], "phone" [m]} with the
    reference's schema (cint_len / pint_len = number of ids WITHOUT the leading GO, as the decoder lengths)."""
    seg = utt["utt_id"]
    seg = seg if isinstance(seg, bytes) else str(seg).encode("utf-8")
    char = np.asarray(utt["char"], np.int64)
    phone = np.asarray(utt.get("phone", np.zeros(0, np.int64)), np.int64)
    logmel = np.asarray(utt["logmel"], np.float32)
    ctx = _enc_map([("segment", _enc_feature("bytes", [seg])),
                    ("logmel_len", _enc_feature("int64", [logmel.shape[0]])),
                    ("cint_len", _enc_feature("int64", [max(len(char) - 1, 0)])),
                    ("pint_len", _enc_feature("int64", [max(len(phone) - 1, 0)]))])

    def flist(kind, rows):
        return b"".join(_field(1, 2, _put_varint(len(f)) + f) for f in (_enc_feature(kind, r) for r in rows))

    lists = _enc_map([("logmel", flist("float", logmel)), ("cint", flist("int64", char[:, None])),
                      ("pint", flist("int64", phone[:, None]))])
    return _field(1, 2, _put_varint(len(ctx)) + ctx) + _field(2, 2, _put_varint(len(lists)) + lists)


def write_tfrecord(path, utterances):
    write_records(path, (make_sequence_example(u) for u in utterances))


# ----------------------------------------------------------------------------- the dataset
class _Iterator(object):
    """make_initializable_iterator(): `initialize()` (re)starts a pass, `get_next()` returns the next padded batch."""

    def __init__(self, dataset):
        self.dataset = dataset
        self._gen = None

    def initialize(self):
        self._gen = self.dataset._batches()

    initializer = property(lambda self: self.initialize)

    def get_next(self):
        if self._gen is None:
            self.initialize()
        try:
            return next(self._gen)
        except StopIteration:
            raise OutOfRangeError("end of sequence")


class SpeechDataset(object):
    """Dataset class for speech datasets (speech_dataset.py:5-60).  params: batch_size, feat_length."""

    SHUFFLE_BUFFER = 4000          # speech_dataset.py:52

    def __init__(self, params, data_files, isTraining, seed=0):
        self.params = params
        self.is_training = isTraining
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.data_set, self.data_iter = self.create_iterator(data_files)

    def get_instance(self, proto):
        """Parse the proto to prepare instance (speech_dataset.py:13-45)."""
        context, lists = parse_sequence_example(proto)
        F = int(self.params.feat_length)
        frames = lists.get("logmel", [])
        logmel = np.stack([v for _, v in frames]).astype(np.float32) if frames else np.zeros((0, F), np.float32)
        if logmel.shape[1] != F:
            raise ValueError("logmel feature width %d != feat_length %d" % (logmel.shape[1], F))

        def ids(name):
            steps = lists.get(name, [])
            return np.array([int(v[0]) for _, v in steps], np.int64)

        def scalar(name):
            return np.int64(context[name][1][0])

        return {"logmel": logmel, "char": ids("cint"), "phone": ids("pint"),
                "logmel_len": scalar("logmel_len"), "char_len": scalar("cint_len"),
                "phone_len": scalar("pint_len"), "utt_id": context["segment"][1][0]}

    def create_iterator(self, data_files):
        """Create iterator for data (speech_dataset.py:47-60)."""
        self.data_files = [data_files] if isinstance(data_files, str) else list(data_files)
        return self, _Iterator(self)

    def _instances(self):
        for path in self.data_files:
            for rec in read_records(path):
                yield self.get_instance(rec)

    def _shuffled(self, it):
        """tf.data shuffle(buffer_size): keep a buffer, emit a random element, refill."""
        buf = []
        for x in it:
            if len(buf) < self.SHUFFLE_BUFFER:
                buf.append(x)
                continue
            j = int(self.rng.integers(len(buf)))
            out, buf[j] = buf[j], x
            yield out
        while buf:
            j = int(self.rng.integers(len(buf)))
            buf[j], buf[-1] = buf[-1], buf[j]
            yield buf.pop()

    @staticmethod
    def padded_batch(items):
        """padded_batch with padded_shapes {logmel: [None, F], char: [None], phone: [None], scalars: []}: zero padding
        to the longest of the batch (speech_dataset.py:53-57)."""
        B = len(items)
        F = items[0]["logmel"].shape[1]
        T = max(x["logmel"].shape[0] for x in items)
        batch = {"logmel": np.zeros((B, T, F), np.float32)}
        for i, x in enumerate(items):
            batch["logmel"][i, :x["logmel"].shape[0]] = x["logmel"]
        for key in ("char", "phone"):
            n = max(len(x[key]) for x in items)
            batch[key] = np.full((B, n), PAD_ID, np.int64)
            for i, x in enumerate(items):
                batch[key][i, :len(x[key])] = x[key]
        for key in ("logmel_len", "char_len", "phone_len"):
            batch[key] = np.array([x[key] for x in items], np.int64)
        batch["utt_id"] = np.array([x["utt_id"] for x in items])
        return batch

    def _batches(self):
        it = self._instances()
        if self.is_training:
            it = self._shuffled(it)
        items, bs = [], int(self.params.batch_size)
        for x in it:
            items.append(x)
            if len(items) == bs:
                yield self.padded_batch(items)
                items = []
        if items:                                   # padded_batch keeps the final partial batch
            yield self.padded_batch(items)

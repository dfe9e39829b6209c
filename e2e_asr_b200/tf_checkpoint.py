"""TensorFlow V2 ("tensor bundle") checkpoints without TensorFlow: read {variable name: ndarray} from
`<prefix>.index` + `<prefix>.data-?????-of-?????`, and write them back.

The reference restores its weights by TF variable name through tf.train.NewCheckpointReader
(tf_utils.py:66-90, beam_search.py:36-47).  This module restates the published on-disk format so those files load
into a VariableStore / BeamSearch by the same names:

  * `<prefix>.index` is a LevelDB-style sorted string table (tensorflow/core/lib/io/table*): data blocks of
    prefix-compressed entries (varint32 shared, non_shared, value_len; key delta; value) followed by a restart
    array, each block trailed by 1 compression byte + masked CRC32C; an index block mapping separator keys to
    BlockHandles (varint64 offset, size); a 48-byte footer (metaindex handle, index handle, padding, magic
    0xdb4775248b80fb57).  Key "" holds a BundleHeaderProto, every other key a BundleEntryProto
    (tensorflow/core/protobuf/tensor_bundle.proto): dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5,
    crc32c = 6 (fixed32), slices = 7.
  * `<prefix>.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset + size).

PARITY UNPINNED: no TensorFlow and no real checkpoint exist in this image, so the format is restated from its
specification and checked by round trips, CRC32C known answers and hand-assembled blocks (tests/test_tf_checkpoint_cpu.py),
not against a file written by TensorFlow.  Only uncompressed blocks (what BundleWriter emits) and unsliced
(non-partitioned) variables are supported; anything else raises.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64,
           10: np.bool_, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ----------------------------------------------------------------------------- CRC32C (Castagnoli), masked as LevelDB
def _make_crc_table():
    table = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _make_crc_table()


def crc32c(data, crc=0):
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(crc):
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(masked):
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- varints / minimal protobuf
def _put_varint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_proto(buf):
    """{field number: [values]} of one message: varints as ints, fixed32/64 as ints, length-delimited as bytes."""
    out, pos = {}, 0
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        field, wire = key >> 3, key & 7
        if wire == 0:
            val, pos = _get_varint(buf, pos)
        elif wire == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = _get_varint(buf, pos)
            val = bytes(buf[pos:pos + n])
            pos += n
        elif wire == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.setdefault(field, []).append(val)
    return out


def _field(num, wire, payload):
    return _put_varint((num << 3) | wire) + payload


def _encode_entry(dtype_id, shape, shard_id, offset, size, crc):
    dims = b"".join(_field(2, 2, _put_varint(len(d)) + d) for d in (_field(1, 0, _put_varint(int(s))) for s in shape))
    msg = _field(1, 0, _put_varint(dtype_id)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if shard_id:
        msg += _field(3, 0, _put_varint(shard_id))
    if offset:
        msg += _field(4, 0, _put_varint(offset))
    msg += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return msg


# ----------------------------------------------------------------------------- table (index file)
def _read_block(buf, offset, size, verify=True):
    raw = buf[offset:offset + size]
    ctype = buf[offset + size]
    if verify:
        stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if unmask_crc(stored) != crc32c(buf[offset:offset + size + 1]):
            raise ValueError("table block at %d: CRC32C mismatch" % offset)
    if ctype != 0:
        raise NotImplementedError("compressed table block (type %d): BundleWriter writes uncompressed blocks" % ctype)
    return raw


def _block_entries(block):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def read_table(path, verify=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("%s is not a table file (bad magic)" % path)
    footer = buf[len(buf) - 48:]
    pos = 0
    _, pos = _get_varint(footer, pos)          # metaindex handle
    _, pos = _get_varint(footer, pos)
    idx_off, pos = _get_varint(footer, pos)
    idx_size, pos = _get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, idx_off, idx_size, verify)):
        off, p = _get_varint(handle, 0)
        size, _ = _get_varint(handle, p)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    return out


def _build_block(entries, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value))
        out += key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, entries, block_entries=64):
    """Writes sorted (key, value) pairs as an uncompressed table file."""
    entries = sorted(entries)
    blob, index = bytearray(), []

    def emit(block):
        off = len(blob)
        blob.extend(block)
        blob.append(0)                                            # kNoCompression
        blob.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return off, len(block)

    for i in range(0, max(len(entries), 1), block_entries):
        chunk = entries[i:i + block_entries]
        off, size = emit(_build_block(chunk))
        last = chunk[-1][0] if chunk else b""
        index.append((last, _put_varint(off) + _put_varint(size)))
    meta_off, meta_size = emit(_build_block([]))
    idx_off, idx_size = emit(_build_block(index, restart_interval=1))
    footer = _put_varint(meta_off) + _put_varint(meta_size) + _put_varint(idx_off) + _put_varint(idx_size)
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    with open(path, "wb") as f:
        f.write(bytes(blob) + footer)


# ----------------------------------------------------------------------------- bundles
def checkpoint_exists(prefix):
    return os.path.exists(prefix + ".index")


def read_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} of a V2 checkpoint `prefix` (all variables, or those in `names`)."""
    entries = read_table(prefix + ".index", verify)
    if not entries or entries[0][0] != b"":
        raise ValueError("%s.index has no bundle header" % prefix)
    header = _parse_proto(entries[0][1])
    num_shards = header.get(1, [1])[0]
    if header.get(2, [0])[0] != 0:
        raise NotImplementedError("big-endian tensor bundle")
    shards = {}
    out = {}
    for key, value in entries[1:]:
        name = key.decode("utf-8")
        if names is not None and name not in names:
            continue
        e = _parse_proto(value)
        if 7 in e:
            raise NotImplementedError("%s: partitioned (sliced) variables are not supported" % name)
        dtype_id = e.get(1, [0])[0]
        if dtype_id not in _DTYPES:
            raise NotImplementedError("%s: tensor dtype enum %d" % (name, dtype_id))
        shape = []
        for shp in e.get(2, []):
            for dim in _parse_proto(shp).get(2, []):
                shape.append(_parse_proto(dim).get(1, [0])[0])
        shard, offset, size = e.get(3, [0])[0], e.get(4, [0])[0], e.get(5, [0])[0]
        if shard not in shards:
            shards[shard] = open("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), "rb").read()
        raw = shards[shard][offset:offset + size]
        dt = np.dtype(_DTYPES[dtype_id])
        if len(raw) != size or size != int(np.prod(shape, dtype=np.int64)) * dt.itemsize:
            raise ValueError("%s: %d bytes for shape %s of %s" % (name, size, shape, dt))
        if verify and 6 in e and unmask_crc(e[6][0]) != crc32c(raw):
            raise ValueError("%s: tensor CRC32C mismatch" % name)
        out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).reshape(shape).astype(dt)
    if names is not None:
        missing = sorted(set(names) - set(out))
        if missing:
            raise KeyError("not in checkpoint %s: %s" % (prefix, missing))
    return out


def write_checkpoint(prefix, tensors):
    """Writes {name: ndarray} as a one-shard V2 checkpoint (the layout BundleWriter produces: tensors in key order,
    back to back in `<prefix>.data-00000-of-00001`)."""
    data, entries = bytearray(), []
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        a = np.asarray(tensors[name], order="C")
        if a.dtype not in _DTYPE_IDS:
            raise NotImplementedError("%s: dtype %s" % (name, a.dtype))
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        entries.append((name.encode("utf-8"),
                        _encode_entry(_DTYPE_IDS[a.dtype], a.shape, 0, len(data), len(raw), mask_crc(crc32c(raw)))))
        data.extend(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda v: _put_varint(len(v)) + v)(_field(1, 0, _put_varint(1))))
    entries.append((b"", header))
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))
    write_table(prefix + ".index", entries)

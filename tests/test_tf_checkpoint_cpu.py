"""TF V2 checkpoint ("tensor bundle") reader / writer without TensorFlow: CRC32C known answers,
the table block format on a hand-assembled block, round trips over many variables (several data blocks, prefix
compression), corruption detection, and loading by TF variable name into the oracle's weight dict."""
import os
import struct

import numpy as np
import pytest

from e2e_asr_b200 import synth, tf_checkpoint as tfc


def test_crc32c_known_answers_and_masking():
    assert tfc.crc32c(b"123456789") == 0xE3069283            # the CRC-32C check value
    assert tfc.crc32c(b"") == 0
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA                # RFC 3720 B.4: 32 zero bytes
    assert tfc.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43       # RFC 3720 B.4: 32 bytes of 0xFF
    assert tfc.crc32c(bytes(range(32))) == 0x46DD794E         # RFC 3720 B.4: ascending
    for c in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert tfc.unmask_crc(tfc.mask_crc(c)) == c and tfc.mask_crc(c) != c


def test_varint_and_block_entries_on_a_hand_assembled_block():
    assert tfc._put_varint(300) == b"\xac\x02" and tfc._get_varint(b"\xac\x02", 0) == (300, 2)
    # entries "apple"->"1", "apply"->"22" (shares "appl"), restart point, "banana"->"" ; restart array [0, 17]; count 2
    blk = bytes([0, 5, 1]) + b"apple" + b"1" + bytes([4, 1, 2]) + b"y" + b"22"
    second = len(blk)
    blk += bytes([0, 6, 0]) + b"banana"
    blk += struct.pack("<III", 0, second, 2)
    assert tfc._block_entries(blk) == [(b"apple", b"1"), (b"apply", b"22"), (b"banana", b"")]


def test_round_trip_many_variables(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"model/layer_%03d/sub/kernel" % i: rng.standard_normal((3, i % 5 + 1)).astype(np.float32)
               for i in range(150)}
    tensors["global_step"] = np.array(1234, np.int64)
    tensors["model/empty"] = np.zeros((0, 4), np.float32)
    tensors["model/double"] = rng.standard_normal((2, 2, 2))
    prefix = str(tmp_path / "ckpt-7")
    tfc.write_checkpoint(prefix, tensors)
    assert tfc.checkpoint_exists(prefix) and os.path.exists(prefix + ".data-00000-of-00001")
    got = tfc.read_checkpoint(prefix)
    assert sorted(got) == sorted(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    sub = tfc.read_checkpoint(prefix, names=["global_step", "model/layer_007/sub/kernel"])
    assert set(sub) == {"global_step", "model/layer_007/sub/kernel"}
    with pytest.raises(KeyError):
        tfc.read_checkpoint(prefix, names=["nope"])
    keys = [k for k, _ in tfc.read_table(prefix + ".index")]
    assert keys == sorted(keys) and keys[0] == b"" and len(keys) == len(tensors) + 1


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "c")
    tfc.write_checkpoint(prefix, {"a/b": np.arange(12, dtype=np.float32).reshape(3, 4)})
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError, match="CRC32C"):
        tfc.read_checkpoint(prefix)
    assert tfc.read_checkpoint(prefix, verify=False)["a/b"].shape == (3, 4)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[3] ^= 0x01
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="CRC32C"):
        tfc.read_table(prefix + ".index")
    with pytest.raises(ValueError, match="magic"):
        open(prefix + ".index", "wb").write(bytes(idx[:-1]) + b"\x00")
        tfc.read_table(prefix + ".index")


def test_model_weights_by_tf_name_through_a_checkpoint(tmp_path):
    """The full synthetic model (TF variable names of SURVEY.md Appendix B) survives the checkpoint bit for bit and the
    oracle computes the same step from it."""
    from oracle import model as om
    cfg = synth.get_config("tiny")
    w = synth.make_weights(cfg, bias_noise=0.1)
    prefix = str(tmp_path / "asr.ckpt-0")
    tfc.write_checkpoint(prefix, dict(w, global_step=np.array(0, np.int64)))
    back = {k: v for k, v in tfc.read_checkpoint(prefix).items() if v.dtype.kind == "f"}
    assert sorted(back) == sorted(w) and all(np.array_equal(back[k], w[k]) for k in w)
    batch = synth.make_batch(cfg)
    a = om.train_step(w, batch, num_layers={"char": 4}, ctc_tasks=cfg.ctc, want_grads=False)
    b = om.train_step(back, batch, num_layers={"char": 4}, ctc_tasks=cfg.ctc, want_grads=False)
    assert a["total_loss"] == b["total_loss"]

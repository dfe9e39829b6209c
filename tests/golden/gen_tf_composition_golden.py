"""tests/golden/tf_composition.npz: outputs of the reference's own graph-building functions -- Seq2SeqModel.get_batch
(seq2seq_model.py:159-197), Encoder._get_pyramid_input (encoder.py:94-119), LossUtils.cross_entropy_loss
(losses.py:6-35), tf_utils.create_shifted_targets (tf_utils.py:4-12) -- EXECUTED with a NumPy-backed stand-in for the
dozen TensorFlow ops they call.  This pins how the reference COMPOSES those ops (zero-padding rule for odd maxima,
ceil-division of lengths, mask / per-example normalisation order, time-major transposes, eval-mode lengths); the ops
themselves follow their documented TF-1.x semantics, stated below.  Run in the build container only."""
import builtins
import contextlib
import importlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from e2e_asr_b200.base_params import Bunch  # noqa: E402


class _Dim(object):
    def __init__(self, v):
        self.value = v


class T(np.ndarray):
    """ndarray with the one Tensor method the reference calls: get_shape()[i].value."""

    def get_shape(self):
        return [_Dim(int(s)) for s in self.shape]


def t(x, dtype=None):
    return np.asarray(x, dtype).view(T)


def make_tf():
    tf = types.ModuleType("tensorflow")
    tf.int32, tf.int64, tf.float32, tf.bool = np.int32, np.int64, np.float32, np.bool_
    tf.concat = lambda values, axis: t(np.concatenate([np.asarray(v) for v in values], axis))
    tf.zeros = lambda shape, dtype=np.float32: t(np.zeros([int(s) for s in shape], dtype))
    tf.shape = lambda x: t(np.array(np.shape(x), np.int32))
    tf.transpose = lambda x, perm: t(np.transpose(x, perm))
    tf.ones_like = lambda x: t(np.ones_like(x))
    tf.reduce_max = lambda x: t(np.max(x))
    tf.mod = lambda a, b: t(np.mod(a, b))
    tf.cast = lambda x, dtype: t(np.asarray(x).astype(dtype))          # float -> int truncates toward zero, as TF
    tf.cond = lambda pred, true_fn, false_fn: true_fn() if bool(pred) else false_fn()
    tf.identity = lambda x: x
    tf.reshape = lambda x, shape: t(np.reshape(x, [int(s) for s in np.asarray(shape).reshape(-1)]))
    tf.to_int64 = lambda x: t(np.asarray(x).astype(np.int64))
    tf.ceil = lambda x: t(np.ceil(x))
    tf.truediv = lambda a, b: t(np.true_divide(a, b))
    tf.slice = lambda x, begin, size: t(x[tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))])

    def sequence_mask(lengths, maxlen=None, dtype=np.bool_):
        lengths = np.asarray(lengths)
        maxlen = int(lengths.max()) if maxlen is None else maxlen      # TF: max(lengths) when maxlen is None
        return t((np.arange(maxlen)[None, :] < lengths[:, None]).astype(dtype))
    tf.sequence_mask = sequence_mask
    tf.reduce_sum = lambda x, reduction_indices=None: t(np.sum(x, axis=reduction_indices))
    tf.reduce_mean = lambda x: t(np.mean(x))
    tf.name_scope = lambda *a, **k: contextlib.nullcontext()
    nn = types.ModuleType("tensorflow.nn")

    def sparse_xent(logits, labels):
        lg = np.asarray(logits, np.float64)
        lse = lg.max(1) + np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1))
        return t((lse - lg[np.arange(lg.shape[0]), np.asarray(labels)]).astype(np.float32))
    nn.sparse_softmax_cross_entropy_with_logits = sparse_xent
    tf.nn = nn
    return tf


def main():
    tf = make_tf()
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.nn"] = tf.nn
    for name in ("tensorflow.contrib", "tensorflow.contrib.rnn", "tensorflow.contrib.rnn.python",
                 "tensorflow.contrib.rnn.python.ops", "tensorflow.contrib.rnn.python.ops.core_rnn_cell"):
        m = types.ModuleType(name)
        m._linear = None
        sys.modules[name] = m
    b = types.ModuleType("bunch")
    b.Bunch = Bunch
    sys.modules["bunch"] = b
    builtins.xrange = range
    sys.path.insert(0, "/root/reference")
    s2s = importlib.import_module("seq2seq_model")
    enc = importlib.import_module("encoder")
    losses = importlib.import_module("losses")
    tfu = importlib.import_module("tf_utils")

    rng = np.random.Generator(np.random.PCG64(42))
    out = {}
    # ---- get_batch: stacking, time-major ids, eval lengths
    B, Tn, F, U = 3, 7, 4, 5
    logmel = rng.standard_normal((B, Tn, F)).astype(np.float32)
    char = rng.integers(0, 9, size=(B, U + 1)).astype(np.int64)
    batch = {"logmel": t(logmel), "logmel_len": t([7, 4, 1], np.int64), "char": t(char),
             "char_len": t([5, 2, 1], np.int64), "utt_id": np.array(["a", "b", "c"])}
    for stack in (1, 3):
        for training in (True, False):
            fake = types.SimpleNamespace(encoder=types.SimpleNamespace(params=Bunch(stack_cons=stack)),
                                         params=Bunch(tasks=["char"], max_output={"char": 11}), isTraining=training)
            enc_in, dec_in, enc_len, dec_len = s2s.Seq2SeqModel.get_batch(fake, batch)
            key = "get_batch/stack%d/%s/" % (stack, "train" if training else "eval")
            out[key + "enc_in"], out[key + "dec_in"] = np.asarray(enc_in), np.asarray(dec_in["char"])
            out[key + "enc_len"], out[key + "dec_len"] = np.asarray(enc_len), np.asarray(dec_len["char"])
    out["get_batch/logmel"], out["get_batch/char"] = logmel, char
    # ---- _get_pyramid_input: odd / even maxima, several lengths
    for name, Tm, lens in (("odd", 7, [7, 4, 1]), ("even", 6, [6, 5, 2]), ("odd_short", 5, [3, 5, 1])):
        x = rng.standard_normal((3, Tm, 4)).astype(np.float32)
        fake = types.SimpleNamespace(params=Bunch(skip_step=2))
        y, new_len = enc.Encoder._get_pyramid_input(fake, t(x), t(lens, np.int64))
        out["pyramid/%s/x" % name], out["pyramid/%s/lens" % name] = x, np.asarray(lens, np.int64)
        out["pyramid/%s/y" % name], out["pyramid/%s/new_len" % name] = np.asarray(y), np.asarray(new_len)
    # ---- cross_entropy_loss and create_shifted_targets
    Uu, Bb, V = 4, 3, 6
    logits = rng.standard_normal((Uu * Bb, V)).astype(np.float32)
    dec_inp = rng.integers(0, V, size=(Uu + 1, Bb)).astype(np.int64)
    seq_len = np.array([4, 2, 1], np.int64)
    targets, weights = tfu.create_shifted_targets(t(dec_inp), t(seq_len))
    loss = losses.LossUtils.cross_entropy_loss(t(logits), targets, t(seq_len))
    out["loss/logits"], out["loss/dec_inp"], out["loss/seq_len"] = logits, dec_inp, seq_len
    out["loss/targets"], out["loss/weights"], out["loss/value"] = np.asarray(targets), np.asarray(weights), np.asarray(loss)
    np.savez(os.path.join(HERE, "tf_composition.npz"), **out)
    return out


if __name__ == "__main__":
    o = main()
    print(len(o), "arrays;", "loss", float(o["loss/value"]), "pyramid odd", o["pyramid/odd/y"].shape, o["pyramid/odd/new_len"])

"""The TensorFlow stand-in of np_tf.py on torch float64 tensors, so that torch.autograd differentiates the reference's
OWN forward graph code (encoder.py, decoder.py, attn_decoder.py, losses.py executed unmodified): the resulting
gradients pin the oracle's hand-derived backward pass (gen_grad_golden.py).  Same op semantics as np_tf.py -- the
scope / naming machinery is imported from it -- with every numeric op expressed in torch.

TEST INFRASTRUCTURE for the golden generators only (build container)."""
import contextlib
import types

import numpy as np
import torch

import np_tf
from np_tf import LSTMStateTuple, _Dim, _Shape, _map_state  # noqa: F401

F64 = torch.float64
_DT = {np.int32: torch.int32, np.int64: torch.int64, np.bool_: torch.bool, np.float64: F64, np.float32: F64}


class T(torch.Tensor):
    def get_shape(self):
        return _Shape(_Dim(int(s)) for s in self.shape)

    def __int__(self):
        return int(self.item())

    __index__ = __int__


def t(x, dtype=None):
    if not isinstance(x, torch.Tensor):
        a = np.asarray(x)
        if a.dtype.kind == "f":
            a = a.astype(np.float64)
        x = torch.from_numpy(np.ascontiguousarray(a)) if a.ndim else torch.tensor(a.item(), dtype=_DT.get(a.dtype.type))
    if dtype is not None:
        x = x.to(_DT.get(dtype, dtype))
    return x.as_subclass(T)


def _ints(shape):
    if isinstance(shape, torch.Tensor):
        return [int(s) for s in shape.reshape(-1)]
    if isinstance(shape, (int, np.integer)):
        return [int(shape)]
    return [int(s) for s in shape]


class Graph(np_tf.Graph):
    def __init__(self, weights):
        np_tf.Graph.__init__(self, weights)
        self.params = {k: torch.tensor(v, dtype=F64, requires_grad=True) for k, v in self.weights.items()}

    def get_variable(self, name, shape=None, dtype=None, initializer=None, **kw):
        full = self.path() + "/" + name if self.scope else name
        if full not in self.params:
            raise KeyError("get_variable(%r): no such weight (scope %r)" % (full, self.path()))
        v = self.params[full]
        if shape is not None:
            assert tuple(_ints(shape)) == tuple(v.shape), (full, shape, tuple(v.shape))
        if full not in self.used:
            self.used.append(full)
        return v.as_subclass(T)


class BasicLSTMCell(np_tf._Cell):
    base_name = "basic_lstm_cell"

    def __init__(self, g, num_units, forget_bias=1.0):
        np_tf._Cell.__init__(self, g)
        self.n, self.forget_bias = num_units, forget_bias

    def zero_state(self, batch_size, dtype=None):
        z = lambda: t(torch.zeros((int(batch_size), self.n), dtype=F64))
        return LSTMStateTuple(z(), z())

    def __call__(self, x, state):
        c, h = state
        with self._enter():
            k = self.g.get_variable("kernel", [x.shape[1] + self.n, 4 * self.n])
            b = self.g.get_variable("bias", [4 * self.n])
        z = torch.cat([x, h], 1) @ k + b
        i, j, f, o = torch.chunk(z, 4, dim=1)
        new_c = c * torch.sigmoid(f + self.forget_bias) + torch.sigmoid(i) * torch.tanh(j)
        new_h = torch.tanh(new_c) * torch.sigmoid(o)
        return t(new_h), LSTMStateTuple(t(new_c), t(new_h))


class GRUCell(np_tf._Cell):
    base_name = "gru_cell"

    def __init__(self, g, num_units):
        np_tf._Cell.__init__(self, g)
        self.n = num_units

    def zero_state(self, batch_size, dtype=None):
        return t(torch.zeros((int(batch_size), self.n), dtype=F64))

    def __call__(self, x, h):
        with self._enter():
            with self.g.variable_scope("gates"):
                gk = self.g.get_variable("kernel", [x.shape[1] + self.n, 2 * self.n])
                gb = self.g.get_variable("bias", [2 * self.n])
            with self.g.variable_scope("candidate"):
                ck = self.g.get_variable("kernel", [x.shape[1] + self.n, self.n])
                cb = self.g.get_variable("bias", [self.n])
        v = torch.sigmoid(torch.cat([x, h], 1) @ gk + gb)
        r, u = v[:, :self.n], v[:, self.n:]
        c = torch.tanh(torch.cat([x, r * h], 1) @ ck + cb)
        new_h = u * h + (1.0 - u) * c
        return t(new_h), t(new_h)


class MultiRNNCell(np_tf.MultiRNNCell):
    pass


def make_tf(weights):
    g = Graph(weights)
    tf = types.ModuleType("tensorflow")
    tf._graph = g
    tf.int32, tf.int64, tf.float32, tf.bool = np.int32, np.int64, np.float64, np.bool_
    tf.AUTO_REUSE = "auto_reuse"
    tf.variable_scope = g.variable_scope
    tf.get_variable = g.get_variable
    tf.random_uniform_initializer = lambda *a, **k: None
    tf.name_scope = lambda *a, **k: contextlib.nullcontext()
    tf.concat = lambda values, axis: t(torch.cat([t(v) for v in values], axis))
    tf.zeros = lambda shape, dtype=np.float64: t(torch.zeros(_ints(shape), dtype=_DT.get(dtype, F64)))
    tf.shape = lambda x: t(torch.tensor(list(x.shape), dtype=torch.int32))
    tf.transpose = lambda x, perm: t(t(x).permute(*perm))
    tf.ones_like = lambda x: t(torch.ones_like(t(x)))
    tf.reduce_max = lambda x: t(torch.max(t(x)))
    tf.reduce_all = lambda x: bool(torch.all(t(x)))
    tf.mod = lambda a, b: t(torch.remainder(t(a), b))
    tf.less = lambda a, b: t(torch.lt(t(a), t(b)))
    tf.cast = lambda x, dtype: t(x).to(_DT[dtype]).as_subclass(T)
    tf.cond = lambda pred, true_fn, false_fn: true_fn() if bool(pred) else false_fn()
    tf.identity = lambda x: x
    tf.reshape = lambda x, shape: t(t(x).reshape(_ints(shape)))
    tf.to_int64 = lambda x: t(x).to(torch.int64).as_subclass(T)
    tf.ceil = lambda x: t(torch.ceil(t(x)))
    tf.truediv = lambda a, b: t(torch.true_divide(t(a), t(b)))
    tf.tanh = lambda x: t(torch.tanh(x))

    def stack(values):
        vs = [t(v) for v in values]
        if not any(v.is_floating_point() for v in vs):
            vs = [v.to(torch.int64) for v in vs]
        return t(torch.stack(vs))
    tf.stack = stack
    tf.tile = lambda x, multiples: t(t(x).repeat(*_ints(multiples)))
    tf.expand_dims = lambda x, axis: t(t(x).unsqueeze(axis))
    tf.argmax = lambda x, axis: t(torch.argmax(t(x), axis))
    tf.slice = lambda x, begin, size: t(x[tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))])

    def reduce_sum(x, axis=None, reduction_indices=None, keepdims=False):
        ax = axis if axis is not None else reduction_indices
        x = t(x)
        return t(torch.sum(x) if ax is None else torch.sum(x, dim=tuple(ax) if isinstance(ax, (list, tuple)) else ax,
                                                            keepdim=keepdims))
    tf.reduce_sum = reduce_sum
    tf.reduce_mean = lambda x: t(torch.mean(t(x)))

    def sequence_mask(lengths, maxlen=None, dtype=np.bool_):
        lengths = t(lengths)
        maxlen = int(lengths.max()) if maxlen is None else maxlen
        return t((torch.arange(maxlen)[None, :] < lengths[:, None]).to(_DT.get(dtype, F64)))
    tf.sequence_mask = sequence_mask

    class TensorArray(object):
        def __init__(self, size=None, dtype=None, **kw):
            self.items = None

        def unstack(self, value):
            ta = TensorArray()
            ta.items = [t(v) for v in value]
            return ta

        def read(self, index):
            return self.items[int(index)]
    tf.TensorArray = TensorArray

    nn = types.ModuleType("tensorflow.nn")
    tf.nn = nn
    nn.embedding_lookup = lambda params, ids: t(params[t(ids).long()])
    nn.softmax = lambda x: t(torch.softmax(t(x), dim=-1))

    def conv2d(x, w, strides, padding):
        assert w.shape[0] == 1 and w.shape[1] == 1 and list(strides) == [1, 1, 1, 1]
        return t(torch.einsum("bthc,cd->bthd", t(x), w[0, 0]))
    nn.conv2d = conv2d

    def sparse_xent(logits, labels):
        lg = t(logits)
        return t(torch.logsumexp(lg, dim=1) - lg[torch.arange(lg.shape[0]), t(labels).long()])
    nn.sparse_softmax_cross_entropy_with_logits = sparse_xent

    rc = types.ModuleType("tensorflow.nn.rnn_cell")
    nn.rnn_cell = rc
    rc.BasicLSTMCell = lambda n, **k: BasicLSTMCell(g, n, **k)
    rc.GRUCell = lambda n, **k: GRUCell(g, n)
    rc.DropoutWrapper = lambda cell, output_keep_prob=1.0, **k: np_tf.DropoutWrapper(cell, output_keep_prob)
    rc.MultiRNNCell = lambda cells, **k: MultiRNNCell(g, cells)
    rc.LSTMStateTuple = LSTMStateTuple

    def _reverse(x, lens):
        out = x.clone()
        for b in range(x.shape[1]):
            n = int(lens[b])
            out[:n, b] = torch.flip(x[:n, b], dims=[0])
        return out

    def dynamic_rnn(cell, inputs, sequence_length=None, dtype=None, time_major=False, scope=None, reverse=False):
        assert time_major
        x = t(inputs)
        Tn, B = x.shape[0], x.shape[1]
        lens = t(sequence_length)
        if reverse:
            x = _reverse(x, lens)
        with g.variable_scope(scope if scope is not None else "rnn"):
            state = cell.zero_state(B, dtype)
            outs = []
            for step in range(Tn):
                out, new_state = cell(t(x[step]), state)
                live = (step < lens)[:, None]
                outs.append(torch.where(live, out, torch.zeros_like(out)))
                state = _map_state(lambda n_, o_: t(torch.where(live, n_, o_)), new_state, state)
        y = torch.stack(outs)
        if reverse:
            y = _reverse(y, lens)
        return t(y), state
    nn.dynamic_rnn = dynamic_rnn

    def bidirectional_dynamic_rnn(cell_fw, cell_bw, inputs, sequence_length=None, dtype=None, time_major=False,
                                  scope=None):
        with g.variable_scope(scope if scope is not None else "bidirectional_rnn"):
            with g.variable_scope("fw") as fw_scope:
                out_fw, st_fw = dynamic_rnn(cell_fw, inputs, sequence_length, dtype, time_major, fw_scope)
            with g.variable_scope("bw") as bw_scope:
                out_bw, st_bw = dynamic_rnn(cell_bw, inputs, sequence_length, dtype, time_major, bw_scope, reverse=True)
        return (out_fw, out_bw), (st_fw, st_bw)
    nn.bidirectional_dynamic_rnn = bidirectional_dynamic_rnn

    def raw_rnn(cell, loop_fn, scope=None):
        with g.variable_scope(scope if scope is not None else "rnn"):
            time = 0
            finished, next_input, state, emit_structure, loop_state = loop_fn(time, None, None, None)
            finished = t(finished)
            emits = []
            while not bool(torch.all(finished)):
                output, cell_state = cell(next_input, state)
                time += 1
                next_finished, next_input, next_state, emit, new_loop_state = loop_fn(time, output, cell_state, loop_state)
                fin = finished[:, None]
                emits.append(torch.where(fin, torch.zeros_like(emit), emit))
                state = _map_state(lambda n_, o_: t(torch.where(fin, o_, n_)), next_state, state)
                if new_loop_state is not None:
                    loop_state = new_loop_state
                finished = torch.logical_or(finished, t(next_finished))

            class _Emit(object):
                def concat(self_inner):
                    return t(torch.cat(emits, dim=0))
        return _Emit(), state, loop_state
    nn.raw_rnn = raw_rnn

    def _linear(args, output_size, bias, **kw):
        if not isinstance(args, (list, tuple)):
            args = [args]
        x = torch.cat([t(a) for a in args], 1)
        k = g.get_variable("kernel", [x.shape[1], int(output_size)])
        y = x @ k
        if bias:
            y = y + g.get_variable("bias", [int(output_size)])
        return t(y)
    tf._linear = _linear
    return tf

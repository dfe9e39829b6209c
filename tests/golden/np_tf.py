"""A NumPy stand-in for the part of TensorFlow 1.x that the reference's graph-building code calls (encoder.py,
decoder.py, attn_decoder.py, losses.py, tf_utils.py, seq2seq_model.get_batch), so that code can be EXECUTED here --
eagerly, on float64 arrays -- to produce golden vectors for the oracle.

TEST INFRASTRUCTURE for the golden generators only (build container; never imported by tests or the product).

What is the reference's and what is restated here:
  * the reference's own Python runs unmodified (one mechanical Python-2 fix at load time: dict.has_key): layer wiring,
    scopes, the order of operations inside raw_loop_function, attention(), the loss composition;
  * the TF ops follow their documented TF-1.x semantics, restated below: element-wise / shape ops, `_linear`,
    BasicLSTMCell / GRUCell / DropoutWrapper / MultiRNNCell, variable_scope / get_variable naming (default_name
    uniquification per parent scope, layers named at first call), dynamic_rnn / bidirectional_dynamic_rnn
    (zero output and state copy-through past sequence_length; bw = reverse_sequence, run, reverse back) and raw_rnn
    (loop_fn protocol, emit zeroing and state copy-through for finished rows, loop_state passed as is).
Variables are not created: get_variable returns the entry of a weight dict keyed by the full TF name (and records
it, so a generator can assert that every weight was consumed under exactly its name)."""
import contextlib
import types

import numpy as np

F64 = np.float64


class _Dim(object):
    """tf.Dimension: .value, and usable where an int is expected."""

    def __init__(self, v):
        self.value = v

    def __int__(self):
        return int(self.value)

    __index__ = __int__


class _Shape(list):
    def with_rank(self, r):
        assert len(self) == r
        return self

    def as_list(self):
        return [d.value for d in self]


class T(np.ndarray):
    def get_shape(self):
        return _Shape(_Dim(int(s)) for s in self.shape)


def t(x, dtype=None):
    return np.asarray(x, dtype).view(T)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class LSTMStateTuple(tuple):
    def __new__(cls, c, h):
        return tuple.__new__(cls, (c, h))

    c = property(lambda self: self[0])
    h = property(lambda self: self[1])


class Graph(object):
    """Variable scopes + the weight dict."""

    def __init__(self, weights):
        self.weights = {k: np.asarray(v, F64) for k, v in weights.items()}
        self.used = []
        self.scope = []                 # list of scope-name components
        self.taken = {}                 # parent path -> set of child names opened with default_name uniquification

    def path(self):
        return "/".join(self.scope)

    @contextlib.contextmanager
    def variable_scope(self, name_or_scope=None, default_name=None, reuse=None, initializer=None, **kw):
        if isinstance(name_or_scope, _ScopeHandle):
            saved, self.scope = self.scope, list(name_or_scope.components)
            try:
                yield name_or_scope
            finally:
                self.scope = saved
            return
        if name_or_scope is None:       # default_name: uniquified among the names taken under this parent
            taken = self.taken.setdefault(self.path(), set())
            name, i = default_name, 0
            while name in taken:
                i += 1
                name = "%s_%d" % (default_name, i)
            taken.add(name)
        else:
            name = name_or_scope
            self.taken.setdefault(self.path(), set()).add(name)
        self.scope = self.scope + name.split("/")
        try:
            yield _ScopeHandle(self.scope)
        finally:
            self.scope = self.scope[:-len(name.split("/"))]

    def get_variable(self, name, shape=None, dtype=None, initializer=None, **kw):
        full = self.path() + "/" + name if self.scope else name
        if full not in self.weights:
            raise KeyError("get_variable(%r): no such weight (scope %r)" % (full, self.path()))
        v = self.weights[full]
        if shape is not None:
            assert tuple(int(s) for s in shape) == v.shape, (full, shape, v.shape)
        if full not in self.used:
            self.used.append(full)
        return t(v)


class _ScopeHandle(object):
    def __init__(self, components):
        self.components = list(components)
        self.name = "/".join(components)


class _Cell(object):
    """RNNCell base: a layer whose variable scope is fixed at its FIRST call (default_name = the class's base name,
    uniquified under the scope of that call)."""
    base_name = None

    def __init__(self, g):
        self.g, self._scope = g, None

    @contextlib.contextmanager
    def _enter(self):
        if self._scope is None:
            with self.g.variable_scope(None, default_name=self.base_name) as sc:
                self._scope = sc
                yield
        else:
            with self.g.variable_scope(self._scope):
                yield


class BasicLSTMCell(_Cell):
    base_name = "basic_lstm_cell"

    def __init__(self, g, num_units, forget_bias=1.0):
        _Cell.__init__(self, g)
        self.n, self.forget_bias = num_units, forget_bias

    def zero_state(self, batch_size, dtype=None):
        return LSTMStateTuple(t(np.zeros((int(batch_size), self.n), F64)), t(np.zeros((int(batch_size), self.n), F64)))

    def __call__(self, x, state):
        c, h = state
        with self._enter():
            k = self.g.get_variable("kernel", [x.shape[1] + self.n, 4 * self.n])
            b = self.g.get_variable("bias", [4 * self.n])
        z = np.concatenate([x, h], 1) @ k + b
        i, j, f, o = np.split(z, 4, axis=1)
        new_c = c * sigmoid(f + self.forget_bias) + sigmoid(i) * np.tanh(j)
        new_h = np.tanh(new_c) * sigmoid(o)
        return t(new_h), LSTMStateTuple(t(new_c), t(new_h))


class GRUCell(_Cell):
    base_name = "gru_cell"

    def __init__(self, g, num_units):
        _Cell.__init__(self, g)
        self.n = num_units

    def zero_state(self, batch_size, dtype=None):
        return t(np.zeros((int(batch_size), self.n), F64))

    def __call__(self, x, h):
        with self._enter():
            with self.g.variable_scope("gates"):
                gk = self.g.get_variable("kernel", [x.shape[1] + self.n, 2 * self.n])
                gb = self.g.get_variable("bias", [2 * self.n])
            with self.g.variable_scope("candidate"):
                ck = self.g.get_variable("kernel", [x.shape[1] + self.n, self.n])
                cb = self.g.get_variable("bias", [self.n])
        v = sigmoid(np.concatenate([x, h], 1) @ gk + gb)
        r, u = v[:, :self.n], v[:, self.n:]
        c = np.tanh(np.concatenate([x, r * h], 1) @ ck + cb)
        new_h = u * h + (1.0 - u) * c
        return t(new_h), t(new_h)


class DropoutWrapper(object):
    """tf.nn.rnn_cell.DropoutWrapper(cell, output_keep_prob=p): the cell's OUTPUT is multiplied by a fresh keep-mask / p
    at every call, the state passes through untouched; p == 1.0 (a Python float) disables it, as in TF.  TF's random
    masks cannot be reproduced, so with p < 1 the generator injects them: hook(wrapper, output) -> mask (already
    scaled by 1/p).  `calls` counts this wrapper's invocations; `ctx` is set by dynamic_rnn before each call (step,
    reverse, lens); `cell._scope.name` is the wrapped cell's variable scope once it has been called."""

    def __init__(self, cell, output_keep_prob=1.0, hook=None, **kw):
        self.cell, self.keep, self.hook, self.calls, self.ctx = cell, float(output_keep_prob), hook, 0, None

    def zero_state(self, *a, **k):
        return self.cell.zero_state(*a, **k)

    def __call__(self, x, state):
        out, new_state = self.cell(x, state)
        if self.keep < 1.0:
            assert self.hook is not None, "dropout masks must be injected (tf._dropout)"
            out = t(out * self.hook(self, out))
        self.calls += 1
        return out, new_state


class MultiRNNCell(_Cell):
    base_name = "multi_rnn_cell"

    def __init__(self, g, cells):
        _Cell.__init__(self, g)
        self.cells = list(cells)

    def zero_state(self, batch_size, dtype=None):
        return tuple(c.zero_state(batch_size, dtype) for c in self.cells)

    def __call__(self, x, state):
        new = []
        with self._enter():
            for i, cell in enumerate(self.cells):
                with self.g.variable_scope("cell_%d" % i):
                    x, s = cell(x, state[i])
                    new.append(s)
        return x, tuple(new)


def _map_state(fn, *states):
    s0 = states[0]
    if isinstance(s0, LSTMStateTuple):
        return LSTMStateTuple(*[_map_state(fn, *[s[i] for s in states]) for i in range(2)])
    if isinstance(s0, tuple):
        return tuple(_map_state(fn, *[s[i] for s in states]) for i in range(len(s0)))
    return fn(*states)


def make_tf(weights):
    g = Graph(weights)
    tf = types.ModuleType("tensorflow")
    tf._graph = g
    tf.int32, tf.int64, tf.float32, tf.bool = np.int32, np.int64, F64, np.bool_      # "float32" computes in float64
    tf.AUTO_REUSE = "auto_reuse"
    tf.variable_scope = g.variable_scope
    tf.get_variable = g.get_variable
    tf.random_uniform_initializer = lambda *a, **k: None
    tf.name_scope = lambda *a, **k: contextlib.nullcontext()
    tf.concat = lambda values, axis: t(np.concatenate([np.asarray(v) for v in values], axis))
    tf.zeros = lambda shape, dtype=F64: t(np.zeros([int(s) for s in np.asarray(shape).reshape(-1)], dtype))
    tf.shape = lambda x: t(np.array(np.shape(x), np.int32))
    tf.transpose = lambda x, perm: t(np.transpose(x, perm))
    tf.ones_like = lambda x: t(np.ones_like(x))
    tf.reduce_max = lambda x: t(np.max(x))
    tf.reduce_all = lambda x: bool(np.all(x))
    tf.mod = lambda a, b: t(np.mod(a, b))
    tf.less = lambda a, b: t(np.less(a, b))
    tf.cast = lambda x, dtype: t(np.asarray(x).astype(dtype))
    tf.cond = lambda pred, true_fn, false_fn: true_fn() if bool(pred) else false_fn()
    tf.identity = lambda x: x
    tf.reshape = lambda x, shape: t(np.reshape(x, [int(s) for s in np.asarray(shape).reshape(-1)]))
    tf.to_int64 = lambda x: t(np.asarray(x).astype(np.int64))
    tf.ceil = lambda x: t(np.ceil(x))
    tf.truediv = lambda a, b: t(np.true_divide(a, b))
    tf.tanh = lambda x: t(np.tanh(x))
    tf.stack = lambda values: t(np.stack([np.asarray(v) for v in values]))
    tf.tile = lambda x, multiples: t(np.tile(x, [int(m) for m in np.asarray(multiples).reshape(-1)]))
    tf.expand_dims = lambda x, axis: t(np.expand_dims(x, axis))
    tf.argmax = lambda x, axis: t(np.argmax(x, axis))
    tf.slice = lambda x, begin, size: t(x[tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))])

    def reduce_sum(x, axis=None, reduction_indices=None, keepdims=False):
        ax = axis if axis is not None else reduction_indices
        if isinstance(ax, list):
            ax = tuple(ax)
        return t(np.sum(x, axis=ax, keepdims=keepdims))
    tf.reduce_sum = reduce_sum
    tf.reduce_mean = lambda x: t(np.mean(x))

    def sequence_mask(lengths, maxlen=None, dtype=np.bool_):
        lengths = np.asarray(lengths)
        maxlen = int(lengths.max()) if maxlen is None else maxlen
        return t((np.arange(maxlen)[None, :] < lengths[:, None]).astype(dtype))
    tf.sequence_mask = sequence_mask

    class TensorArray(object):
        def __init__(self, size=None, dtype=None, **kw):
            self.items = None

        def unstack(self, value):
            ta = TensorArray()
            ta.items = [t(v) for v in np.asarray(value)]
            return ta

        def read(self, index):
            return self.items[int(index)]
    tf.TensorArray = TensorArray

    nn = types.ModuleType("tensorflow.nn")
    tf.nn = nn
    nn.embedding_lookup = lambda params, ids: t(np.asarray(params)[np.asarray(ids)])

    def softmax(x):
        e = np.exp(x - np.max(x, axis=-1, keepdims=True))
        return t(e / e.sum(axis=-1, keepdims=True))
    nn.softmax = softmax

    def conv2d(x, w, strides, padding):
        assert w.shape[0] == 1 and w.shape[1] == 1 and list(strides) == [1, 1, 1, 1]
        return t(np.einsum("bthc,cd->bthd", x, np.asarray(w)[0, 0]))
    nn.conv2d = conv2d

    def sparse_xent(logits, labels):
        lg = np.asarray(logits, F64)
        lse = lg.max(1) + np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1))
        return t(lse - lg[np.arange(lg.shape[0]), np.asarray(labels)])
    nn.sparse_softmax_cross_entropy_with_logits = sparse_xent

    rc = types.ModuleType("tensorflow.nn.rnn_cell")
    nn.rnn_cell = rc
    rc.BasicLSTMCell = lambda n, **k: BasicLSTMCell(g, n, **k)
    rc.GRUCell = lambda n, **k: GRUCell(g, n)
    tf._dropout = None
    rc.DropoutWrapper = lambda cell, output_keep_prob=1.0, **k: DropoutWrapper(
        cell, output_keep_prob, hook=lambda wr, out: tf._dropout(wr, out))
    rc.MultiRNNCell = lambda cells, **k: MultiRNNCell(g, cells)
    rc.LSTMStateTuple = LSTMStateTuple

    def dynamic_rnn(cell, inputs, sequence_length=None, dtype=None, time_major=False, scope=None, reverse=False):
        assert time_major
        x = np.asarray(inputs, F64)
        Tn, B = x.shape[0], x.shape[1]
        lens = np.asarray(sequence_length)
        if reverse:                                   # array_ops.reverse_sequence(seq_axis=0, batch_axis=1)
            xr = x.copy()
            for b in range(B):
                xr[:lens[b], b] = x[:lens[b], b][::-1]
            x = xr
        with g.variable_scope(scope if scope is not None else "rnn"):
            state = cell.zero_state(B, dtype)
            outs = []
            for step in range(Tn):
                if isinstance(cell, DropoutWrapper):
                    cell.ctx = dict(step=step, reverse=reverse, lens=lens)
                out, new_state = cell(t(x[step]), state)
                live = (step < lens)[:, None]
                outs.append(np.where(live, out, 0.0))                               # zero output past the length
                state = _map_state(lambda n_, o_: t(np.where(live, n_, o_)), new_state, state)   # copy-through
        y = np.stack(outs)
        if reverse:
            yr = y.copy()
            for b in range(B):
                yr[:lens[b], b] = y[:lens[b], b][::-1]
            y = yr
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
            finished = np.asarray(finished)
            emits = []
            while not np.all(finished):
                output, cell_state = cell(next_input, state)
                time += 1
                next_finished, next_input, next_state, emit, new_loop_state = loop_fn(time, output, cell_state, loop_state)
                fin = finished[:, None]
                emits.append(np.where(fin, 0.0, emit))                                       # zero_emit for finished rows
                state = _map_state(lambda n_, o_: t(np.where(fin, o_, n_)), next_state, state)   # copy state through
                if new_loop_state is not None:
                    loop_state = new_loop_state
                finished = np.logical_or(finished, np.asarray(next_finished))

            class _Emit(object):
                def concat(self_inner):
                    return t(np.concatenate(emits, axis=0))

                def stack(self_inner):
                    return t(np.stack(emits))
        return _Emit(), state, loop_state
    nn.raw_rnn = raw_rnn

    def _linear(args, output_size, bias, **kw):
        if not isinstance(args, (list, tuple)):
            args = [args]
        x = np.concatenate([np.asarray(a, F64) for a in args], 1)
        k = g.get_variable("kernel", [x.shape[1], int(output_size)])
        y = x @ k
        if bias:
            y = y + g.get_variable("bias", [int(output_size)])
        return t(y)
    tf._linear = _linear

    # Training-graph plumbing of Seq2SeqModel.__init__ / create_computational_graph (seq2seq_model.py:74-157): state
    # variables, summaries, the optimiser and tf.gradients are inert here -- only the forward composition is executed.
    class _Var(object):
        def __init__(self, value, trainable=True, **kw):
            self.value = value

        def assign(self, v):
            return None

        def __mul__(self, o):
            return self.value * o

        def __add__(self, o):
            return self.value + o
    tf.Variable = _Var
    tf.summary = types.SimpleNamespace(scalar=lambda *a, **k: None, merge_all=lambda: None)
    tf.trainable_variables = lambda: []
    tf.gradients = lambda loss, variables: []
    tf.clip_by_global_norm = lambda grads, clip: ([], 0.0)
    tf.train = types.SimpleNamespace(AdamOptimizer=lambda lr: types.SimpleNamespace(
        apply_gradients=lambda gv, global_step=None: None))

    # Randomness of scheduled sampling (attn_decoder.py:131-136, decoder.py:176): TF's generators cannot be reproduced,
    # so a generator may inject draws -- tf._draws = dict(uniform=fn(step) -> scalar, multinomial=fn(step, logits) ->
    # ids); tf.random_uniform([]) is evaluated once per loop step (step = number of calls so far), tf.multinomial uses
    # the step of the preceding uniform draw.
    tf._draws, tf._step = None, [0]

    def random_uniform(shape, *a, **k):
        assert list(shape) == [] and tf._draws is not None
        tf._step[0] += 1
        return t(tf._draws["uniform"](tf._step[0]))
    tf.random_uniform = random_uniform

    def multinomial(logits, num_samples, *a, **k):
        assert num_samples == 1 and tf._draws is not None
        return t(np.asarray(tf._draws["multinomial"](tf._step[0], np.asarray(logits))).reshape(-1, 1))
    tf.multinomial = multinomial
    return tf

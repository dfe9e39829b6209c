"""Host-side (NumPy) utilities that keep the reference's names and signatures:
`BasicLSTM` (reference basic_lstm.py:10-23), `sigmoid` / `softmax`
(num_utils.py:6-14) and the `BeamEntry` hypothesis record (beam_entry.py:1-23).
They are used by the float64 beam-search scorer and by user code; the training
path never touches them.
"""
import collections

import numpy as np


def sigmoid(x):
    """Logistic function, elementwise."""
    return np.reciprocal(1.0 + np.exp(np.negative(x)))


def softmax(x):
    """Softmax of a 1-D score vector, shifted by its maximum for stability."""
    shifted = np.exp(x - x.max())
    return shifted / shifted.sum(axis=0)


class BasicLSTM(object):
    """Single-vector TF BasicLSTMCell: gates in i, j, f, o column order, forget
    bias of 1 added at run time, state returned as (c, h)."""

    FORGET_BIAS = 1.0

    def __init__(self, weight, bias):
        self.lstm_w, self.lstm_b = weight, bias

    def __call__(self, x, lstm_state):
        c_prev, h_prev = lstm_state
        pre = np.concatenate((x, h_prev), axis=0) @ self.lstm_w + self.lstm_b
        in_gate, cand, forget, out_gate = np.split(pre, 4)
        c_next = c_prev * sigmoid(forget + self.FORGET_BIAS) + sigmoid(in_gate) * np.tanh(cand)
        return (c_next, sigmoid(out_gate) * np.tanh(c_next))


_Entry = collections.namedtuple("_Entry", ["index_seq", "dec_state", "context_vec", "cum_attn_probs"])


class BeamEntry(_Entry):
    """Immutable beam hypothesis: token ids so far, decoder states, context vector."""
    __slots__ = ()

    def __new__(cls, index_seq, dec_state, context_vec, cum_attn_probs=None):
        return super(BeamEntry, cls).__new__(cls, index_seq, dec_state, context_vec, cum_attn_probs)

    def get_last_output(self):
        return self.index_seq[-1]

    get_index_seq = property(lambda self: (lambda: self.index_seq))
    get_dec_state = property(lambda self: (lambda: self.dec_state))
    get_context_vec = property(lambda self: (lambda: self.context_vec))
    get_cum_attn_probs = property(lambda self: (lambda: self.cum_attn_probs))


def philox_uniform(counter, offset, seed):
    """Word 0 of Philox4x32-10(counter=(counter, offset, 0, 0), key=(seed lo, seed hi)) * 2^-32: the host-side scalar
    draw of the same counter-based generator the CUDA dropout / sampling kernels use (csrc/misc.cu)."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c0, c1, c2, c3 = counter & MASK, offset & MASK, 0, 0
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0 * 2.0 ** -32

"""Deterministic synthetic batches and weights.

The reference ships no data, vocabularies or checkpoints, so every parity and
benchmark run uses the generator below: numpy PCG64 streams keyed by seed,
weights drawn with the reference's initialisers (encoder kernels U(+-0.075),
reference encoder.py:73-74; embedding U(+-1), decoder.py:97-99; every other
kernel Glorot-uniform, the TF default when no initialiser is in scope; biases 0)
and keyed by the TF variable names of SURVEY.md Appendix B, which are the
reference's weight interchange format (beam_search.py:56-98).

Pure numpy: used by the product path (bench, smoke) and by the tests/oracle.
"""
import math

import numpy as np

from .base_params import Bunch
from .data_utils import EOS_ID, GO_ID, PAD_ID  # noqa: F401

DATA_SEED = 1234
WEIGHT_SEED = 4321

# name -> (B, T, F, H, L, V, U); aux CTC heads as in SURVEY.md section 8d.
CONFIGS = {
    # tiny: unit-test size, oracle float64 in milliseconds
    "tiny": dict(B=3, T=22, F=8, H=8, L=4, V=17, U=6, E=8, A=8, Hd=8, Hl=8,
                 ctc={"phone_ctc": (3, 5), "state": (2, 9)}),
    "tiny_b": dict(B=5, T=37, F=12, H=16, L=4, V=23, U=9, E=12, A=8, Hd=16, Hl=8,
                   ctc={"phone_ctc": (3, 6)}),
    # tiny_uni / uni256: forward-only encoder (bi_dir=False, tf.nn.dynamic_rnn)
    "tiny_uni": dict(B=4, T=29, F=10, H=16, L=3, V=19, U=7, E=12, A=8, Hd=16, Hl=8, bi_dir=False,
                     ctc={"phone_ctc": (2, 6)}),
    "uni256": dict(B=3, T=40, F=16, H=256, L=2, V=31, U=6, E=32, A=16, Hd=32, Hl=32, bi_dir=False,
                   ctc={"phone_ctc": (2, 6)}),
    # tiny_gru / gru256: GRU encoder cells (the reference's Encoder.class_params() default, encoder.py:27,48)
    "tiny_gru": dict(B=5, T=31, F=10, H=16, L=3, V=19, U=7, E=12, A=8, Hd=16, Hl=8, enc_lstm=False,
                     ctc={"phone_ctc": (2, 6)}),
    "gru256": dict(B=6, T=44, F=16, H=256, L=2, V=31, U=6, E=32, A=16, Hd=32, Hl=32, enc_lstm=False,
                   ctc={"phone_ctc": (2, 6)}),
    # decoder cell variants (decoder.py:49-82): stacked cells (MultiRNNCell) and GRU cells in lm_cell + decoder cell
    "tiny_dec2": dict(B=4, T=26, F=10, H=16, L=3, V=19, U=7, E=12, A=8, Hd=16, Hl=8, dec_layers=2,
                      ctc={"phone_ctc": (2, 6)}),
    "tiny_decgru": dict(B=4, T=26, F=10, H=16, L=3, V=19, U=7, E=12, A=8, Hd=16, Hl=8, dec_lstm=False,
                        ctc={"phone_ctc": (2, 6)}),
    "tiny_decgru2": dict(B=4, T=26, F=10, H=16, L=3, V=19, U=7, E=12, A=8, Hd=16, Hl=16, dec_layers=2,
                         dec_lstm=False, ctc={}),
    # wide_small: cfg-5's widths at unit-test size (H=512 -> L2-exchange recurrence, D=1024 -> per-step decoder)
    "wide_small": dict(B=3, T=24, F=8, H=512, L=2, V=17, U=5, E=16, A=16, Hd=32, Hl=16,
                       ctc={"phone_ctc": (1, 5)}),
    # cfg-1: base_params defaults, CPU parity
    "cfg1": dict(B=4, T=200, F=40, H=256, L=4, V=1000, U=25, E=256, A=128,
                 Hd=256, Hl=256, ctc={"phone_ctc": (3, 48)}),
    # cfg-2: Switchboard-300h-shaped training batch (the bench workload)
    "cfg2": dict(B=64, T=700, F=120, H=256, L=4, V=1000, U=120, E=256, A=128,
                 Hd=256, Hl=256, ctc={"phone_ctc": (3, 48), "state": (2, 1024)}),
    # cfg-4: long-utterance stress
    "cfg4": dict(B=32, T=2000, F=120, H=256, L=4, V=1000, U=120, E=256, A=128,
                 Hd=256, Hl=256, ctc={"phone_ctc": (3, 48), "state": (2, 1024)}),
    # cfg-5: wide encoder
    "cfg5": dict(B=256, T=700, F=120, H=512, L=5, V=1000, U=120, E=256, A=128,
                 Hd=256, Hl=256, ctc={"phone_ctc": (3, 48), "state": (2, 1024)}),
}


def get_config(name, **overrides):
    cfg = Bunch(CONFIGS[name])
    cfg.name = name
    cfg.ctc = dict(cfg.ctc)
    for k, v in overrides.items():
        cfg[k] = v
    return cfg


def layer_input_sizes(cfg, skip_step=None, max_scaling_down=None):
    """Input width of every encoder layer (reference encoder.py:154-178); cfg may carry the encoder options
    skip_step / max_scaling_down / initial_res_fac (cfg.F is the width AFTER frame stacking)."""
    skip_step = cfg.get("skip_step", 2) if skip_step is None else skip_step
    max_scaling_down = cfg.get("max_scaling_down", 8) if max_scaling_down is None else max_scaling_down
    sizes, res = [], cfg.get("initial_res_fac", 1)
    width = cfg.F
    nd = 2 if cfg.get("bi_dir", True) else 1
    for i in range(cfg.L):
        sizes.append(width)
        if skip_step > 1 and i != cfg.L - 1 and res < max_scaling_down:
            width = nd * cfg.H * skip_step
            res *= skip_step
        else:
            width = nd * cfg.H
    return sizes


def _glorot(rng, shape, fan_in=None, fan_out=None):
    if fan_in is None:
        if len(shape) == 1:
            fan_in = fan_out = shape[0]
        else:
            recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            fan_in, fan_out = shape[-2] * recept, shape[-1] * recept
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def make_weights(cfg, seed=WEIGHT_SEED, tasks=("char",), bias_noise=0.0):
    """Weights keyed by TF variable name.

    `bias_noise` > 0 perturbs the (reference: zero-initialised) biases so parity
    tests exercise the bias paths with non-trivial values.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    H = cfg.H

    def bias(n):
        if bias_noise > 0:
            return rng.uniform(-bias_noise, bias_noise, size=(n,)).astype(np.float32)
        return np.zeros((n,), np.float32)

    bi_dir = cfg.get("bi_dir", True)
    for l, I in enumerate(layer_input_sizes(cfg), start=1):
        scopes = ["model/encoder/RNNLayer%d/bidirectional_rnn/%s/" % (l, d) for d in ("fw", "bw")] \
            if bi_dir else ["model/encoder/RNNLayer%d/%d/" % (l, l)]
        for sc in scopes:
            if cfg.get("enc_lstm", True):
                w[sc + "basic_lstm_cell/kernel"] = rng.uniform(-0.075, 0.075, size=(I + H, 4 * H)).astype(np.float32)
                w[sc + "basic_lstm_cell/bias"] = bias(4 * H)
            else:   # GRUCell: gates bias initialised to 1.0, candidate bias to 0 (tf.nn.rnn_cell.GRUCell)
                w[sc + "gru_cell/gates/kernel"] = rng.uniform(-0.075, 0.075, size=(I + H, 2 * H)).astype(np.float32)
                w[sc + "gru_cell/gates/bias"] = (1.0 + bias(2 * H)).astype(np.float32)
                w[sc + "gru_cell/candidate/kernel"] = rng.uniform(-0.075, 0.075, size=(I + H, H)).astype(np.float32)
                w[sc + "gru_cell/candidate/bias"] = bias(H)
    D = (2 if bi_dir else 1) * H
    for task in tasks:
        V = cfg.V if task == "char" else cfg.get("V_" + task, cfg.V)
        p = "model/rnn_decoder_%s/" % task
        w[p + "decoder/embedding"] = rng.uniform(-1.0, 1.0, size=(V, cfg.E)).astype(np.float32)
        w[p + "AttnW"] = _glorot(rng, (1, 1, D, cfg.A))
        w[p + "AttnV"] = _glorot(rng, (cfg.A,))
        # lm_cell (called first in the raw_rnn loop) and the decoder cell: single cells or MultiRNNCell stacks of
        # LSTM / GRU cells (decoder.py:49-72); names as in oracle _cell_names
        nl, lstm = cfg.get("dec_layers", 1), cfg.get("dec_lstm", True)
        for which, (sfx, Hc) in enumerate((("", cfg.Hl), ("_1", cfg.Hd))):
            for l in range(nl):
                cell = "basic_lstm_cell" if lstm else "gru_cell"
                base = p + "rnn/" + (cell + sfx + "/" if nl == 1 else "multi_rnn_cell%s/cell_%d/%s/" % (sfx, l, cell))
                I_ = cfg.E if l == 0 else Hc
                if lstm:
                    w[base + "kernel"] = _glorot(rng, (I_ + Hc, 4 * Hc))
                    w[base + "bias"] = bias(4 * Hc)
                else:
                    w[base + "gates/kernel"] = _glorot(rng, (I_ + Hc, 2 * Hc))
                    w[base + "gates/bias"] = (1.0 + bias(2 * Hc)).astype(np.float32)
                    w[base + "candidate/kernel"] = _glorot(rng, (I_ + Hc, Hc))
                    w[base + "candidate/bias"] = bias(Hc)
        w[p + "rnn/Attention/kernel"] = _glorot(rng, (cfg.Hd, cfg.A))
        w[p + "rnn/Attention/bias"] = bias(cfg.A)
        w[p + "rnn/AttnProjection/kernel"] = _glorot(rng, (cfg.Hd + D, cfg.Hd))
        w[p + "rnn/AttnProjection/bias"] = bias(cfg.Hd)
        w[p + "rnn/OutputProjection/kernel"] = _glorot(rng, (cfg.Hd, V))
        w[p + "rnn/OutputProjection/bias"] = bias(V)
        w[p + "rnn/InputProjection/kernel"] = _glorot(rng, (cfg.Hd + D, cfg.E))
        w[p + "rnn/InputProjection/bias"] = bias(cfg.E)
        if cfg.Hl != cfg.Hd:
            w[p + "rnn/SimpleProjection/kernel"] = _glorot(rng, (cfg.Hl, cfg.Hd))
            w[p + "rnn/SimpleProjection/bias"] = bias(cfg.Hd)
    for task, (depth, vocab) in cfg.ctc.items():
        p = "model/ctc_%s/" % task
        w[p + "kernel"] = _glorot(rng, (D, vocab + 1))
        w[p + "bias"] = bias(vocab + 1)
    return w


def pyramid_lens(lens, n):
    """Sequence lengths after n pyramid reductions (reference encoder.py:117-118)."""
    lens = np.asarray(lens, np.int64)
    for _ in range(n):
        lens = (lens + 1) // 2
    return lens


def depth_reductions(cfg, depth, skip_step=2, max_scaling_down=8):
    """Number of x2 reductions applied before the output of layer `depth`."""
    n, res = 0, 1
    for i in range(depth - 1):
        if skip_step > 1 and res < max_scaling_down:
            n += 1
            res *= skip_step
    return n


def make_batch(cfg, seed=DATA_SEED, tasks=("char",)):
    """One padded batch with the layout of reference speech_dataset.py:43-45."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, T, F, U = cfg.B, cfg.T, cfg.F, cfg.U
    lens = rng.integers(max(1, int(math.ceil(0.6 * T))), T + 1, size=B).astype(np.int64)
    lens[rng.integers(0, B)] = T
    logmel = rng.standard_normal(size=(B, T, F)).astype(np.float32)
    for b in range(B):
        logmel[b, lens[b]:] = 0.0
    batch = {"logmel": logmel, "logmel_len": lens,
             "utt_id": np.array(["utt%05d" % i for i in range(B)])}
    for task in tasks:
        V = cfg.V if task == "char" else cfg.get("V_" + task, cfg.V)
        tl = rng.integers(max(1, U // 2), U + 1, size=B).astype(np.int64)
        tl[rng.integers(0, B)] = U
        ids = np.full((B, U + 1), PAD_ID, np.int64)
        for b in range(B):
            n = int(tl[b])  # number of targets, includes the trailing EOS
            ids[b, 0] = GO_ID
            if n > 1:
                ids[b, 1:n] = rng.integers(3, V, size=n - 1)
            ids[b, n] = EOS_ID
        batch[task] = ids
        batch[task + "_len"] = tl
    for task, (depth, vocab) in cfg.ctc.items():
        dl = pyramid_lens(lens, depth_reductions(cfg, depth))
        ll = np.array([rng.integers(1, max(2, int(dl[b]) // 2 + 1)) for b in range(B)], np.int64)
        lab = np.zeros((B, int(ll.max())), np.int64)
        for b in range(B):
            lab[b, :ll[b]] = rng.integers(0, vocab, size=int(ll[b]))
        batch[task] = lab
        batch[task + "_len"] = ll
    return batch


def make_beam_eval_batch(cfg, n_utts=256, seed=17):
    """BASELINE.json configs[2]: a synthetic eval batch of encoder top-layer states, T_enc uniform in [50, 88]
    (what eval_model.py:141-143 hands to BeamSearch one utterance at a time)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return [(np.tanh(rng.standard_normal((int(rng.integers(50, 89)), 2 * cfg.H))) * 0.8).astype(np.float32)
            for _ in range(n_utts)]

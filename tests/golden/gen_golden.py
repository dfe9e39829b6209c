"""Generate golden vectors by EXECUTING the reference's own NumPy code.

Run in the build container only (`python tests/golden/gen_golden.py`); it reads
/root/reference, which does not exist on the GPU box.  The committed *.npz files
are what the tests use.

What runs unmodified from /root/reference: basic_lstm.py, num_utils.py,
beam_entry.py, beam_search.py (BeamSearch.__init__, map_*_variables,
calc_attention, top_k_setup_with_lm/get_top_k, __call__).  The reference is
Python 2 / TF1 code, so three shims restore its environment -- none touches the
arithmetic:
  * `tensorflow` and `bunch` are stubbed modules (tf is only used to read a
    checkpoint; `tf_utils.get_matching_variables` is replaced by a function that
    returns our synthetic weights keyed by the same TF variable names);
  * `xrange` is bound to `range`;
  * beam_search.py's module-level `np` is a proxy restoring Python-2 / old-numpy
    semantics at exactly two call sites: `np.zeros(shape/4)` where `/` was
    integer division (beam_search.py:236-243), and
    `np.divide(idx, k, dtype=np.int32)` which floor-divided ints (:306).
"""
import builtins
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from e2e_asr_b200 import synth  # noqa: E402
from e2e_asr_b200.base_params import Bunch  # noqa: E402


class _NpPy2(object):
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def zeros(shape, *a, **k):
        if isinstance(shape, float):
            assert shape == int(shape)
            shape = int(shape)
        return np.zeros(shape, *a, **k)

    @staticmethod
    def divide(a, b, dtype=None, **k):
        if dtype is not None and np.issubdtype(dtype, np.integer):
            return np.floor_divide(a, b).astype(dtype)
        return np.divide(a, b, **k)


def load_reference():
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    bunch_mod = types.ModuleType("bunch")
    bunch_mod.Bunch = Bunch
    sys.modules["bunch"] = bunch_mod
    builtins.xrange = range
    sys.path.insert(0, REF)
    import basic_lstm, num_utils, beam_search, tf_utils  # noqa: E401
    beam_search.np = _NpPy2()
    return basic_lstm, num_utils, beam_search, tf_utils


def dec_weights(cfg, seed, eos_bias, out_scale=1.0):
    """Synthetic decoder weights.  An untrained decoder has near-uniform output
    distributions, and without length normalisation (beam_search.py:336) the
    empty hypothesis would always win; sharpening the output projection and
    biasing EOS gives a mix of immediate-EOS, mid-length and 120-step outputs."""
    w = synth.make_weights(cfg, seed=seed, bias_noise=0.05)
    w = {k: v for k, v in w.items() if "rnn_decoder_char" in k}
    w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] *= np.float32(out_scale)
    w["model/rnn_decoder_char/rnn/OutputProjection/bias"][2] += np.float32(eos_bias)
    return w


BEAM_CASES = [
    # (tag, config, weight seed, out scale, eos bias, n_utt, T range, beam sizes)
    ("tiny", "tiny", 11, 14.0, 2.0, 8, (3, 20), (1, 2, 4)),
    ("tinyb", "tiny_b", 12, 8.0, 2.0, 8, (4, 20), (1, 4, 10)),   # Hl != Hd: SimpleProjection path
    ("tinyb2", "tiny_b", 13, 25.0, 4.0, 8, (4, 20), (1, 4, 10)),
    ("cfg1", "cfg1", 13, 14.0, 3.0, 4, (20, 40), (1, 4, 10)),
]


def main():
    basic_lstm, num_utils, beam_search, tf_utils = load_reference()
    rng = np.random.Generator(np.random.PCG64(99))

    # --- cell / numerics known answers (basic_lstm.py, num_utils.py) ---
    out = {}
    for i, (I, H) in enumerate([(5, 4), (16, 8), (40, 32)]):
        w = rng.uniform(-0.5, 0.5, size=(I + H, 4 * H)).astype(np.float32)
        b = rng.uniform(-0.5, 0.5, size=(4 * H,)).astype(np.float32)
        x = rng.standard_normal(I)
        c = rng.standard_normal(H)
        h = np.tanh(rng.standard_normal(H))
        nc, nh = basic_lstm.BasicLSTM(w, b)(x, (c, h))
        out.update({"lstm%d_w" % i: w, "lstm%d_b" % i: b, "lstm%d_x" % i: x, "lstm%d_c" % i: c,
                    "lstm%d_h" % i: h, "lstm%d_nc" % i: nc, "lstm%d_nh" % i: nh})
    v = rng.standard_normal(37) * 4
    out["num_x"] = v
    out["num_sigmoid"] = num_utils.sigmoid(v)
    out["num_softmax"] = num_utils.softmax(v)
    np.savez_compressed(os.path.join(HERE, "cell_numerics.npz"), **out)

    # --- decoder step + beam search (beam_search.py) ---
    for tag, cname, wseed, out_scale, eos_bias, n_utt, (tlo, thi), beams in BEAM_CASES:
        cfg = synth.get_config(cname)
        w = dec_weights(cfg, wseed, eos_bias, out_scale)
        tf_utils.get_matching_variables = lambda substr, path, _w=w: dict(_w)
        res = {"weight_seed": wseed, "eos_bias": eos_bias, "out_scale": out_scale, "config": cname,
               "beams": np.array(beams), "n_utt": n_utt}
        encs = []
        for u in range(n_utt):
            T = int(rng.integers(tlo, thi + 1))
            encs.append((np.tanh(rng.standard_normal((T, 2 * cfg.H))) * 0.8).astype(np.float32))
            res["enc%d" % u] = encs[-1]
        for k in beams:
            sp = beam_search.BeamSearch.class_params()
            sp.beam_size = k
            sp.lm_path = "same-checkpoint"
            bs = beam_search.BeamSearch("synthetic-checkpoint", sp)
            for u, enc in enumerate(encs):
                # 2-D [T_enc, D] as eval_model.py:141 passes it.  (A [1,T,D] input only
                # works in the reference when T == D: __call__ sizes its zero
                # context from shape[1] (:246) before calc_attention squeezes.)
                ids = bs(enc)
                res["ids_k%d_u%d" % (k, u)] = np.asarray(ids, np.int64)
        # one raw decoder step (get_top_k) from GO with zero states: full score vector
        sp = beam_search.BeamSearch.class_params()
        sp.beam_size = 3
        sp.lm_weight = 0.3
        sp.lm_path = "same-checkpoint"
        bs = beam_search.BeamSearch("synthetic-checkpoint", sp)
        fn = bs.top_k_setup_with_lm(encs[0])
        hs = w["model/rnn_decoder_char/rnn/basic_lstm_cell_1/kernel"].shape[1] // 4
        ls = w["model/rnn_decoder_char/rnn/basic_lstm_cell/kernel"].shape[1] // 4
        z = lambda n: (np.zeros(n), np.zeros(n))
        x = bs.dec_params.embedding[1]
        top, ms, ts, st, ctx = fn(x, x, [z(hs), z(ls), z(ls)], np.zeros(encs[0].shape[1]), beam_size=3)
        x2 = bs.dec_params.embedding[int(top[0])]
        top2, ms2, ts2, st2, ctx2 = fn(x2, x2, st, ctx, beam_size=3)
        res.update(step_top=np.sort(top), step_scores=np.sort(ms), step_ctx=ctx,
                   step_dec_c=st[0][0], step_dec_h=st[0][1], step_lm_h=st[1][1],
                   step2_in=int(top[0]), step2_top=np.sort(top2), step2_scores=np.sort(ms2), step2_ctx=ctx2,
                   step2_dec_c=st2[0][0])
        np.savez_compressed(os.path.join(HERE, "beam_%s.npz" % tag), **res)
        print("wrote beam_%s.npz" % tag, {k: [len(res["ids_k%d_u%d" % (k, u)]) for u in range(n_utt)] for k in beams})


def gen_relevant_words():
    """tests/golden/relevant_words.json: the reference's own data_utils.get_relevant_words (data_utils.py:20-33,
    `tensorflow` stubbed) on sample transcripts -- pins e2e_asr_b200/scoring.py."""
    import importlib
    import json
    sys.modules.setdefault("tensorflow", types.ModuleType("tensorflow"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    du = importlib.import_module("data_utils")
    samples = ["hello<sp>world", "uh i mean um the the-  cat", "[noise] yeah [laughter] right-", "", "  a   b  ",
               "ha-ha haha ha", "well<sp>uh<sp>y-<sp>you know", "mm hm okay", "eee-<sp>ew<sp>er<sp>error",
               "[vocalized-noise]<sp>so<sp>-"]
    out = [{"in": s_, "words": du.get_relevant_words(s_)[0], "rel": du.get_relevant_words(s_)[1]} for s_ in samples]
    with open(os.path.join(HERE, "relevant_words.json"), "w") as f:
        json.dump(out, f, indent=0)


if __name__ == "__main__":
    gen_relevant_words()
    main()

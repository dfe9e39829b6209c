"""tests/golden/graph_*.npz: encoder states, decoder logits and losses produced by EXECUTING the reference's own
Encoder.__call__ (encoder.py:122-180), AttnDecoder.__call__ (attn_decoder.py:37-172, incl. raw_loop_function and
attention()), Decoder.prepare_decoder_input / get_cell / get_state (decoder.py) and LossUtils.cross_entropy_loss on
the NumPy stand-in for TensorFlow of np_tf.py, with the synthetic weights of e2e_asr_b200.synth keyed by TF variable
name.  The generator also asserts that the reference consumed EVERY weight under exactly the name synth / the product
give it (scope structure of the reference + TF's naming rules as restated in np_tf.py).

Configurations: bidirectional LSTM encoder + single-LSTM decoder (the benchmarked model), forward-only encoder, GRU
encoder, 2-layer LSTM decoder, GRU decoder, 2-layer GRU decoder.  Run in the build container only; the reference
sources are copied to a temporary directory with one mechanical Python-2 fix (dict.has_key -> in)."""
import builtins
import importlib
import os
import re
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import np_tf  # noqa: E402
from e2e_asr_b200 import synth  # noqa: E402
from e2e_asr_b200.base_params import Bunch  # noqa: E402

REF = "/root/reference"
CASES = {   # name -> (synth config, encoder params, decoder params)
    "lstm": ("tiny_b", {}, {}),
    "uni": ("tiny_uni", {"bi_dir": False}, {}),
    "gru_enc": ("tiny_gru", {"use_lstm": False}, {}),
    "dec2": ("tiny_dec2", {}, {"num_layers_dec": 2}),
    "decgru": ("tiny_decgru", {}, {"use_lstm": False}),
    "decgru2": ("tiny_decgru2", {}, {"num_layers_dec": 2, "use_lstm": False}),
}


def load_reference(tf):
    tmp = tempfile.mkdtemp(prefix="ref_py3_")
    for name in ("encoder", "decoder", "attn_decoder", "losses", "tf_utils", "base_params", "seq2seq_model",
                 "data_utils"):
        src = open(os.path.join(REF, name + ".py")).read()
        src = re.sub(r"(\w+)\.has_key\(([^)]*)\)", r"(\2 in \1)", src)
        open(os.path.join(tmp, name + ".py"), "w").write(src)
    sys.modules["tensorflow"] = tf
    for name in ("tensorflow.contrib", "tensorflow.contrib.rnn", "tensorflow.contrib.rnn.python",
                 "tensorflow.contrib.rnn.python.ops", "tensorflow.contrib.rnn.python.ops.core_rnn_cell"):
        m = types.ModuleType(name)
        m._linear = tf._linear
        sys.modules[name] = m
    b = types.ModuleType("bunch")
    b.Bunch = Bunch
    sys.modules["bunch"] = b
    builtins.xrange = range
    sys.path.insert(0, tmp)
    mods = {}
    for name in ("base_params", "data_utils", "encoder", "decoder", "attn_decoder", "losses", "tf_utils",
                 "seq2seq_model"):
        sys.modules.pop(name, None)
    for name in ("base_params", "encoder", "decoder", "attn_decoder", "losses", "tf_utils", "seq2seq_model"):
        mods[name] = importlib.import_module(name)
    sys.path.remove(tmp)
    shutil.rmtree(tmp)
    return mods


def run_decoder_modes_general():
    """graph_modes_general.npz: eval mode and scheduled sampling (as run_decoder_modes) for the decoder's OTHER cell
    configurations (decoder.py:49-72): a 2-layer MultiRNNCell of LSTM cells and GRU cells."""
    from oracle import model as om
    out = {}
    for cname in ("tiny_dec2", "tiny_decgru"):
        cfg = synth.get_config(cname)
        w = synth.make_weights(cfg, bias_noise=0.1)
        w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] = w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] * 6.0
        batch = synth.make_batch(cfg)
        W64 = {k: v.astype(np.float64) for k, v in w.items()}
        states, lens_d, _ = om.encoder_fwd(W64, batch["logmel"].astype(np.float64), batch["logmel_len"], {"char": cfg.L})
        for mode in ("eval", "sample"):
            tf = np_tf.make_tf(w)
            mods = load_reference(tf)
            dp = mods["attn_decoder"].AttnDecoder.class_params()
            dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
            dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
            dp.num_layers_dec, dp.use_lstm = int(cfg.get("dec_layers", 1)), bool(cfg.get("dec_lstm", True))
            dp.out_prob_dec, dp.samp_prob = 1.0, 0.5 if mode == "sample" else 0.0
            seed, task_index, B = 91, 0, cfg.B
            if mode == "sample":
                w0 = lambda step: float(om.philox4x32_10(np.array([step], np.uint64), np.array([200 + task_index], np.uint64),
                                                         np.zeros(1, np.uint64), np.zeros(1, np.uint64), seed & 0xFFFFFFFF,
                                                         (seed >> 32) & 0xFFFFFFFF)[0][0]) * 2.0 ** -32
                tf._draws = dict(uniform=w0, multinomial=lambda step, lg: om.sample_rows(lg, seed, 300 + task_index, step * B))
            seq_len = batch["char_len"] if mode == "sample" else np.full(cfg.B, cfg.U, np.int64)
            with tf.variable_scope("model"):
                dec = mods["attn_decoder"].AttnDecoder(isTraining=(mode == "sample"), params=dp, scope="char")
                logits = dec(np_tf.t(np.ascontiguousarray(batch["char"].T)), np_tf.t(seq_len), np_tf.t(states[cfg.L]),
                             np_tf.t(lens_d[cfg.L]))
            out["%s/%s/logits" % (cname, mode)] = np.asarray(logits)
    out["seed"] = np.array(91)
    np.savez(os.path.join(HERE, "graph_modes_general.npz"), **out)
    return out


def run_decoder_modes():
    """graph_modes.npz: the reference decoder in eval mode (greedy feedback through _get_argmax, lengths = max_output,
    seq2seq_model.py:191-193) and in training mode with scheduled sampling (samp_prob = 0.5; the scalar uniform draw
    per step and the multinomial draw are injected from the oracle's Philox streams, so the reference's own branching
    -- tf.less(random_prob, 1 - samp_prob), finished rows -- decides which ids are fed)."""
    from oracle import model as om
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] = w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] * 6.0
    batch = synth.make_batch(cfg)
    W64 = {k: v.astype(np.float64) for k, v in w.items()}
    states, lens_d, _ = om.encoder_fwd(W64, batch["logmel"].astype(np.float64), batch["logmel_len"], {"char": cfg.L})
    out = {}
    for mode in ("eval", "sample"):
        tf = np_tf.make_tf(w)
        mods = load_reference(tf)
        dp = mods["attn_decoder"].AttnDecoder.class_params()
        dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
        dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
        dp.out_prob_dec, dp.samp_prob = 1.0, 0.5 if mode == "sample" else 0.0
        seed, task_index, B = 77, 0, cfg.B
        if mode == "sample":
            use = om.sample_decisions(cfg.U + 1, 0.5, seed, task_index)
            w0 = lambda step: float(om.philox4x32_10(np.array([step], np.uint64), np.array([200 + task_index], np.uint64),
                                                     np.zeros(1, np.uint64), np.zeros(1, np.uint64), seed & 0xFFFFFFFF,
                                                     (seed >> 32) & 0xFFFFFFFF)[0][0]) * 2.0 ** -32
            tf._draws = dict(uniform=w0, multinomial=lambda step, lg: om.sample_rows(lg, seed, 300 + task_index, step * B))
            out["sample/use"] = use
        seq_len = batch["char_len"] if mode == "sample" else np.full(cfg.B, cfg.U, np.int64)
        with tf.variable_scope("model"):
            dec = mods["attn_decoder"].AttnDecoder(isTraining=(mode == "sample"), params=dp, scope="char")
            logits = dec(np_tf.t(np.ascontiguousarray(batch["char"].T)), np_tf.t(seq_len), np_tf.t(states[cfg.L]),
                         np_tf.t(lens_d[cfg.L]))
        out[mode + "/logits"] = np.asarray(logits)
    out["seed"] = np.array(77)
    np.savez(os.path.join(HERE, "graph_modes.npz"), **out)
    return out


def run_multitask():
    """graph_multitask.npz: Seq2SeqModel.__init__ + create_computational_graph (seq2seq_model.py:50-144) executed for
    the reference's multitask setup -- a char decoder on the top layer and a phone decoder one layer below (main.py:
    89-93), frame stacking on -- with avg = True and False: per-task losses and total_loss."""
    cfg = synth.get_config("tiny_b", V_phone=13)
    base_F = cfg.F
    cfg.F = base_F * 2                       # weights for stack_cons = 2
    tasks = ("char", "phone")
    w = synth.make_weights(cfg, tasks=tasks, bias_noise=0.1)
    cfg.F = base_F
    batch = synth.make_batch(cfg, tasks=tasks)
    out = {}
    for avg in (True, False):
        tf = np_tf.make_tf(w)
        mods = load_reference(tf)
        S = mods["seq2seq_model"].Seq2SeqModel
        p = S.class_params()
        p.tasks, p.num_layers, p.max_output, p.avg = list(tasks), {"char": cfg.L, "phone": cfg.L - 1}, \
            {"char": cfg.U, "phone": cfg.U}, avg
        p.encoder_params.hidden_size, p.encoder_params.use_lstm, p.encoder_params.out_prob = cfg.H, True, 1.0
        p.encoder_params.stack_cons = 2
        dps = {}
        for task in tasks:
            dp = mods["attn_decoder"].AttnDecoder.class_params()
            dp.hidden_size_dec, dp.emb_size = cfg.Hd, cfg.E
            dp.vocab_size = cfg.V if task == "char" else 13
            dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
            dp.out_prob_dec, dp.samp_prob = 1.0, 0.0
            dps[task] = dp
        p.decoder_params = dps
        feed = {k: (np_tf.t(v.astype(np.float64)) if k == "logmel" else (np_tf.t(v) if v.dtype.kind in "if" else v))
                for k, v in batch.items()}
        it = types.SimpleNamespace(get_next=lambda: feed)
        with tf.variable_scope("model"):
            model = S(it, isTraining=True, params=p)
        key = "avg" if avg else "sum"
        for task in tasks:
            out["%s/loss/%s" % (key, task)] = np.asarray(model.losses[task])
            out["%s/logits/%s" % (key, task)] = np.asarray(model.outputs[task])
        out["%s/total_loss" % key] = np.asarray(model.total_loss)
        expect = {k for k in w if not k.startswith("model/ctc_")}         # (the CTC heads are not the reference's)
        assert set(tf._graph.used) == expect, sorted(expect ^ set(tf._graph.used))
    np.savez(os.path.join(HERE, "graph_multitask.npz"), **out)
    return out


def run_dropout():
    """graph_dropout.npz: the reference graph in training mode with out_prob = 0.8 and out_prob_dec = 0.7
    (DropoutWrapper on every encoder cell, on lm_cell and on the decoder cell: encoder.py:49-52, decoder.py:60-63), the
    oracle's Philox keep-masks injected at the positions the oracle uses: encoder layer l -> mask [B, Tp, 2H] of
    stream l (fw | bw halves, the bw cell sees reversed time); lm_cell -> mask [U, B, Hl] of stream 100 + task.  The
    DECODER cell's output is multiplied by ZERO: the reference never reads it (raw_loop_function takes the cell state,
    attn_decoder.py:114-118), so the result must not change -- which is why the oracle applies no mask there."""
    from oracle import model as om
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    keep_e, keep_d, seed = 0.8, 0.7, 31
    tf = np_tf.make_tf(w)
    mods = load_reference(tf)
    B, T = batch["logmel"].shape[:2]
    Tp = om.padded_len(T, {"char": cfg.L})
    U = int(batch["char_len"].max())

    def inject(wr, out):
        scope = wr.cell._scope.name
        if "/encoder/" in scope:
            layer = int(scope.split("RNNLayer")[1].split("/")[0])
            tp_l = Tp >> (layer - 1)
            C = 2 * cfg.H
            mk = om.dropout_mask(B * tp_l * C, keep_e, seed, layer).reshape(B, tp_l, C)
            half = slice(cfg.H, 2 * cfg.H) if "/bw/" in scope else slice(0, cfg.H)
            step, lens = wr.ctx["step"], wr.ctx["lens"]
            tt = np.where(step < lens, lens - 1 - step, 0) if wr.ctx["reverse"] else np.full(B, step)
            return mk[np.arange(B), tt][:, half]
        if scope.endswith("rnn/basic_lstm_cell"):                       # lm_cell: one call per decoder step
            mk = om.dropout_mask(U * B * cfg.Hl, keep_d, seed, 100).reshape(U, B, cfg.Hl)
            return mk[min(wr.calls, U - 1)]
        assert scope.endswith("rnn/basic_lstm_cell_1")                  # the decoder cell: its output is never read
        return np.zeros_like(out)
    tf._dropout = inject
    ep = mods["encoder"].Encoder.class_params()
    ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, True, keep_e
    dp = mods["attn_decoder"].AttnDecoder.class_params()
    dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
    dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
    dp.out_prob_dec, dp.samp_prob = keep_d, 0.0
    with tf.variable_scope("model"):
        enc = mods["encoder"].Encoder(params=ep, isTraining=True)
        att, _, lens = enc(np_tf.t(batch["logmel"].astype(np.float64)), np_tf.t(batch["logmel_len"]), {"char": cfg.L})
        dec = mods["attn_decoder"].AttnDecoder(isTraining=True, params=dp, scope="char")
        dec_inp, seq_len = np_tf.t(np.ascontiguousarray(batch["char"].T)), np_tf.t(batch["char_len"])
        logits = dec(dec_inp, seq_len, att[cfg.L], lens[cfg.L])
        targets, _ = mods["tf_utils"].create_shifted_targets(dec_inp, seq_len)
        loss = mods["losses"].LossUtils.cross_entropy_loss(logits, targets, seq_len)
    out = {"states": np.asarray(att[cfg.L]), "logits": np.asarray(logits), "loss": np.asarray(loss),
           "seed": np.array(seed), "keep": np.array([keep_e, keep_d])}
    np.savez(os.path.join(HERE, "graph_dropout.npz"), **out)
    return out


def run_dropout_stacked():
    """graph_dropout_dec2.npz: a 2-layer LSTM decoder (MultiRNNCell of DropoutWrapper-ed cells, decoder.py:53-68) with
    out_prob_dec = 0.7: the oracle's masks injected on every lm layer's output (stream 100 + task for the top layer,
    500 + 16 task + l below it) and on the lower decoder layers' outputs (400 + 16 task + l); ZEROS on the top decoder
    layer's output, which the reference never reads."""
    from oracle import model as om
    cfg = synth.get_config("tiny_dec2")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    keep, seed, L = 0.7, 13, 2
    tf = np_tf.make_tf(w)
    mods = load_reference(tf)
    B = batch["logmel"].shape[0]
    U = int(batch["char_len"].max())

    def inject(wr, out):
        scope = wr.cell._scope.name
        layer = int(scope.split("/cell_")[1].split("/")[0])
        if "/multi_rnn_cell/" in scope:                                  # lm_cell stack
            stream = 100 if layer == L - 1 else 500 + layer
            return om.dropout_mask(U * B * cfg.Hl, keep, seed, stream).reshape(U, B, cfg.Hl)[min(wr.calls, U - 1)]
        assert "/multi_rnn_cell_1/" in scope                             # decoder cell stack
        if layer == L - 1:
            return np.zeros_like(out)
        return om.dropout_mask(U * B * cfg.Hd, keep, seed, 400 + layer).reshape(U, B, cfg.Hd)[min(wr.calls, U - 1)]
    tf._dropout = inject
    ep = mods["encoder"].Encoder.class_params()
    ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, True, 1.0
    dp = mods["attn_decoder"].AttnDecoder.class_params()
    dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
    dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
    dp.out_prob_dec, dp.samp_prob, dp.num_layers_dec = keep, 0.0, L
    with tf.variable_scope("model"):
        enc = mods["encoder"].Encoder(params=ep, isTraining=True)
        att, _, lens = enc(np_tf.t(batch["logmel"].astype(np.float64)), np_tf.t(batch["logmel_len"]), {"char": cfg.L})
        dec = mods["attn_decoder"].AttnDecoder(isTraining=True, params=dp, scope="char")
        dec_inp, seq_len = np_tf.t(np.ascontiguousarray(batch["char"].T)), np_tf.t(batch["char_len"])
        logits = dec(dec_inp, seq_len, att[cfg.L], lens[cfg.L])
        targets, _ = mods["tf_utils"].create_shifted_targets(dec_inp, seq_len)
        loss = mods["losses"].LossUtils.cross_entropy_loss(logits, targets, seq_len)
    out = {"logits": np.asarray(logits), "loss": np.asarray(loss), "seed": np.array(seed), "keep": np.array(keep)}
    np.savez(os.path.join(HERE, "graph_dropout_dec2.npz"), **out)
    return out


def run_cfg1():
    """graph_cfg1.npz: the reference graph at its default widths (cfg-1: B=4, T=200, F=40, H=Hd=Hl=256, A=128, V=1000,
    4 pyramid layers) -- loss, and every 97th logit / 89th top-layer state value (the full tensors would be megabytes)."""
    cfg = synth.get_config("cfg1")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    tf = np_tf.make_tf(w)
    mods = load_reference(tf)
    ep = mods["encoder"].Encoder.class_params()
    ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, True, 1.0
    dp = mods["attn_decoder"].AttnDecoder.class_params()
    dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
    dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
    dp.out_prob_dec, dp.samp_prob = 1.0, 0.0
    with tf.variable_scope("model"):
        enc = mods["encoder"].Encoder(params=ep, isTraining=True)
        att, _, lens = enc(np_tf.t(batch["logmel"].astype(np.float64)), np_tf.t(batch["logmel_len"]), {"char": cfg.L})
        dec = mods["attn_decoder"].AttnDecoder(isTraining=True, params=dp, scope="char")
        dec_inp, seq_len = np_tf.t(np.ascontiguousarray(batch["char"].T)), np_tf.t(batch["char_len"])
        logits = dec(dec_inp, seq_len, att[cfg.L], lens[cfg.L])
        targets, _ = mods["tf_utils"].create_shifted_targets(dec_inp, seq_len)
        loss = mods["losses"].LossUtils.cross_entropy_loss(logits, targets, seq_len)
    out = {"loss": np.asarray(loss), "logits_shape": np.array(logits.shape),
           "logits_sub": np.asarray(logits).reshape(-1)[::97].copy(),
           "states_sub": np.asarray(att[cfg.L]).reshape(-1)[::89].copy(), "lens": np.asarray(lens[cfg.L])}
    np.savez(os.path.join(HERE, "graph_cfg1.npz"), **out)
    return out


ENC_OPTS = {"res2": dict(initial_res_fac=2), "noskip": dict(skip_step=1),
            "res3_down2": dict(initial_res_fac=3, max_scaling_down=2), "down4": dict(max_scaling_down=4)}


def run_encoder_options():
    """graph_encopts.npz: Encoder.__call__ with the input-stride / pyramid options (encoder.py:149-153,170-176):
    initial_res_fac, skip_step = 1, max_scaling_down; states and lengths at every depth."""
    out = {}
    for name, opts in ENC_OPTS.items():
        cfg = synth.get_config("tiny_b", ctc={}, **opts)
        w = synth.make_weights(cfg, bias_noise=0.1)
        batch = synth.make_batch(cfg)
        tf = np_tf.make_tf(w)
        mods = load_reference(tf)
        ep = mods["encoder"].Encoder.class_params()
        ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, True, 1.0
        ep.update(opts)
        with tf.variable_scope("model"):
            enc = mods["encoder"].Encoder(params=ep, isTraining=True)
            att, _, lens = enc(np_tf.t(batch["logmel"].astype(np.float64)), np_tf.t(batch["logmel_len"]),
                               {"t%d" % d: d for d in range(1, cfg.L + 1)})
        for d in att:
            out["%s/states/%d" % (name, d)] = np.asarray(att[d])
            out["%s/lens/%d" % (name, d)] = np.asarray(lens[d])
    np.savez(os.path.join(HERE, "graph_encopts.npz"), **out)
    return out


def run_case(case):
    cname, enc_over, dec_over = CASES[case]
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    tf = np_tf.make_tf(w)
    mods = load_reference(tf)
    ep = mods["encoder"].Encoder.class_params()
    ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, True, 1.0
    ep.update(enc_over)
    dp = mods["attn_decoder"].AttnDecoder.class_params()
    dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
    dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
    dp.out_prob_dec, dp.samp_prob = 1.0, 0.0
    dp.update(dec_over)
    out = {}
    with tf.variable_scope("model"):
        enc = mods["encoder"].Encoder(params=ep, isTraining=True)
        att, tm, lens = enc(np_tf.t(batch["logmel"].astype(np.float64)), np_tf.t(batch["logmel_len"]),
                            {"char": cfg.L, "state": max(1, cfg.L - 1)})
        for d, v in att.items():
            out["states/%d" % d] = np.asarray(v)
        for d, v in tm.items():
            out["time_major/%d" % d] = np.asarray(v)
        for d, v in lens.items():
            out["lens/%d" % d] = np.asarray(v)
        dec = mods["attn_decoder"].AttnDecoder(isTraining=True, params=dp, scope="char")
        dec_inp = np_tf.t(np.ascontiguousarray(batch["char"].T))
        seq_len = np_tf.t(batch["char_len"])
        logits = dec(dec_inp, seq_len, att[cfg.L], lens[cfg.L])
        targets, weights = mods["tf_utils"].create_shifted_targets(dec_inp, seq_len)
        loss = mods["losses"].LossUtils.cross_entropy_loss(logits, targets, seq_len)
    out["logits"], out["loss"] = np.asarray(logits), np.asarray(loss)
    used = set(tf._graph.used)
    expect = {k for k in w if not k.startswith("model/ctc_")}
    assert used == expect, (sorted(expect - used), sorted(used - expect))
    np.savez(os.path.join(HERE, "graph_%s.npz" % case), **out)
    return out, len(used)


if __name__ == "__main__":
    o = run_decoder_modes()
    print("modes", {k: v.shape for k, v in o.items()})
    o = run_decoder_modes_general()
    print("modes (general cells)", {k: v.shape for k, v in o.items()})
    o = run_cfg1()
    print("cfg1", "loss", float(o["loss"]), o["logits_shape"])
    o = run_dropout()
    print("dropout", "loss", float(o["loss"]))
    o = run_dropout_stacked()
    print("dropout, 2-layer decoder", "loss", float(o["loss"]))
    o = run_encoder_options()
    print("encoder options", sorted({k.split("/")[0] for k in o}), len(o), "arrays")
    o = run_multitask()
    print("multitask", {k: float(v) for k, v in o.items() if "loss" in k})
    for case in CASES:
        o, n = run_case(case)
        print(case, "loss", float(o["loss"]), "logits", o["logits"].shape, "variables consumed", n)

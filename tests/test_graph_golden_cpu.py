"""The oracle's forward pass against golden vectors produced by EXECUTING the reference's own Encoder.__call__,
AttnDecoder.__call__ (raw_loop_function, attention()), Decoder.get_cell / get_state / prepare_decoder_input and
LossUtils.cross_entropy_loss on a NumPy stand-in for TensorFlow (tests/golden/gen_graph_golden.py, np_tf.py), for the
benchmarked model and for every cell variant: forward-only encoder, GRU encoder, 2-layer LSTM decoder, GRU decoder,
2-layer GRU decoder.  (The generator also asserted that the reference consumed every weight under exactly the TF
variable name synth.make_weights / the product's classes give it.)"""
import os

import numpy as np
import pytest

from e2e_asr_b200 import synth
from oracle import model as om

CASES = {"lstm": ("tiny_b", {}, None), "uni": ("tiny_uni", {"bi_dir": False}, None),
         "gru_enc": ("tiny_gru", {"use_lstm": False}, None),
         "dec2": ("tiny_dec2", {}, {"num_layers_dec": 2, "use_lstm": True}),
         "decgru": ("tiny_decgru", {}, {"num_layers_dec": 1, "use_lstm": False}),
         "decgru2": ("tiny_decgru2", {}, {"num_layers_dec": 2, "use_lstm": False})}


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_forward_equals_the_executed_reference_graph(case):
    cname, enc_params, dec_params = CASES[case]
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_%s.npz" % case))
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    res = om.train_step(w, batch, num_layers={"char": cfg.L, "state": max(1, cfg.L - 1)}, tasks=("char",),
                        ctc_tasks={}, enc_params=enc_params or None, dec_params=dec_params, want_grads=False)
    for key in G.files:
        if key.startswith("states/"):
            d = int(key.split("/")[1])
            np.testing.assert_allclose(res["states"][d], G[key], rtol=0, atol=1e-13)
        elif key.startswith("time_major/"):                 # the "state" task's view: same values, time-major
            d = int(key.split("/")[1])
            np.testing.assert_allclose(res["states"][d].transpose(1, 0, 2), G[key], rtol=0, atol=1e-13)
        elif key.startswith("lens/"):
            assert np.array_equal(res["lens"][int(key.split("/")[1])], G[key])
    np.testing.assert_allclose(res["logits"]["char"], G["logits"], rtol=0, atol=1e-12)
    assert abs(res["losses"]["char"] - float(G["loss"])) < 1e-12


def _decoder_inputs():
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] = w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] * 6.0
    batch = synth.make_batch(cfg)
    W64 = {k: v.astype(np.float64) for k, v in w.items()}
    states, lens_d, _ = om.encoder_fwd(W64, batch["logmel"].astype(np.float64), batch["logmel_len"], {"char": cfg.L})
    return cfg, W64, batch, states[cfg.L], lens_d[cfg.L]


def test_greedy_mode_equals_the_executed_reference_eval_graph():
    """isTraining=False: the reference's loop function embeds argmax(previous logits) (decoder.py:139-153,
    attn_decoder.py:128-129) and runs max_output steps for every row (seq2seq_model.py:191-193)."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_modes.npz"))
    cfg, W, batch, enc, enc_len = _decoder_inputs()
    logits, _ = om.attn_decoder_fwd(W, "char", batch["char"].T, np.full(cfg.B, cfg.U), enc, enc_len, mode="greedy",
                                    max_steps=cfg.U)
    np.testing.assert_allclose(logits, G["eval/logits"], rtol=0, atol=1e-11)
    assert np.array_equal(logits.reshape(cfg.U, cfg.B, -1).argmax(2), G["eval/logits"].reshape(cfg.U, cfg.B, -1).argmax(2))


def test_scheduled_sampling_equals_the_executed_reference_training_graph():
    """samp_prob = 0.5: the reference's own branching (one scalar draw per step, tf.less(random_prob, 1 - samp_prob),
    multinomial over the previous logits; attn_decoder.py:130-139, decoder.py:155-180) executed with the oracle's
    Philox draws injected for tf.random_uniform / tf.multinomial."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_modes.npz"))
    cfg, W, batch, enc, enc_len = _decoder_inputs()
    seed = int(G["seed"])
    logits, cache = om.attn_decoder_fwd(W, "char", batch["char"].T, batch["char_len"], enc, enc_len, mode="sample",
                                        samp=(0.5, seed, 0))
    assert G["sample/use"][1:logits.shape[0] // cfg.B].any()                # some steps really sampled
    teacher = batch["char"].T[:len(cache["toks"])]
    assert (np.stack(cache["toks"]) != teacher).any()
    np.testing.assert_allclose(logits, G["sample/logits"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("cname", ["tiny_dec2", "tiny_decgru"])
def test_general_cell_decoders_eval_and_sampling_equal_the_executed_reference(cname):
    """The decoder's other cell configurations (2-layer MultiRNNCell, GRU cells; decoder.py:49-72) in eval mode (greedy
    feedback) and with scheduled sampling (samp_prob = 0.5, the oracle's Philox draws injected): the reference's own
    graph code executed on the NumPy stand-in."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_modes_general.npz"))
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] = w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] * 6.0
    batch = synth.make_batch(cfg)
    W = {k: v.astype(np.float64) for k, v in w.items()}
    states, lens_d, _ = om.encoder_fwd(W, batch["logmel"].astype(np.float64), batch["logmel_len"], {"char": cfg.L})
    enc, enc_len = states[cfg.L], lens_d[cfg.L]
    kw = dict(num_layers_dec=int(cfg.get("dec_layers", 1)), use_lstm=bool(cfg.get("dec_lstm", True)))
    logits, _ = om.attn_decoder_general(W, "char", batch["char"].T, np.full(cfg.B, cfg.U), enc, enc_len, mode="greedy",
                                        max_steps=cfg.U, **kw)
    np.testing.assert_allclose(logits, G[cname + "/eval/logits"], rtol=0, atol=1e-11)
    logits, bwd = om.attn_decoder_general(W, "char", batch["char"].T, batch["char_len"], enc, enc_len, mode="sample",
                                          samp=(0.5, int(G["seed"]), 0), **kw)
    assert (bwd.toks != batch["char"].T[:len(bwd.toks)]).any()              # some inputs really were sampled
    np.testing.assert_allclose(logits, G[cname + "/sample/logits"], rtol=0, atol=1e-11)


def test_multitask_model_equals_the_executed_reference_seq2seq_graph():
    """Seq2SeqModel.__init__ + create_computational_graph (seq2seq_model.py:50-144) executed for char + phone attention
    decoders on different encoder layers with frame stacking: per-task logits / losses and total_loss (avg on / off)."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_multitask.npz"))
    cfg = synth.get_config("tiny_b", V_phone=13)
    base_F = cfg.F
    cfg.F = base_F * 2
    tasks = ("char", "phone")
    w = synth.make_weights(cfg, tasks=tasks, bias_noise=0.1)
    cfg.F = base_F
    batch = synth.make_batch(cfg, tasks=tasks)
    for avg, key in ((True, "avg"), (False, "sum")):
        res = om.train_step(w, batch, tasks=tasks, num_layers={"char": cfg.L, "phone": cfg.L - 1}, ctc_tasks={}, avg=avg,
                            enc_params={"stack_cons": 2}, want_grads=False)
        for task in tasks:
            np.testing.assert_allclose(res["logits"][task], G["%s/logits/%s" % (key, task)], rtol=0, atol=1e-12)
            assert abs(res["losses"][task] - float(G["%s/loss/%s" % (key, task)])) < 1e-12
        assert abs(res["total_loss"] - float(G["%s/total_loss" % key])) < 1e-12
    assert abs(float(G["sum/total_loss"]) - 2 * float(G["avg/total_loss"])) < 1e-12


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_backward_equals_autograd_through_the_executed_reference_graph(case):
    """The oracle's hand-derived backward pass against torch.autograd differentiating the reference's OWN forward code
    (encoder.py / decoder.py / attn_decoder.py / losses.py executed on the torch-backed TensorFlow stand-in,
    tests/golden/gen_grad_golden.py) -- what tf.gradients(total_loss, trainable_vars) computes (seq2seq_model.py:148) --
    for every weight of every cell configuration."""
    cname, enc_params, dec_params = CASES[case]
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "grad_%s.npz" % case))
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    res = om.train_step(w, batch, num_layers={"char": cfg.L}, tasks=("char",), ctc_tasks={}, max_gradient_norm=1e9,
                        enc_params=enc_params or None, dec_params=dec_params)
    assert abs(res["total_loss"] - float(G["loss"])) < 1e-12
    names = [k[len("grad/"):] for k in G.files if k.startswith("grad/")]
    assert sorted(names) == sorted(res["grads"])
    gmax = max(float(np.abs(G["grad/" + k]).max()) for k in names)
    for k in names:
        np.testing.assert_allclose(res["grads"][k], G["grad/" + k], rtol=0, atol=1e-11 * max(1.0, gmax), err_msg=k)


def test_encoder_options_equal_the_executed_reference_encoder():
    """initial_res_fac (input stride + ceil lengths), skip_step = 1 (no pyramid), max_scaling_down (where the pyramid
    stops): encoder.py:149-153,170-176 executed; states and lengths at every depth."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_encopts.npz"))
    opts = {"res2": dict(initial_res_fac=2), "noskip": dict(skip_step=1),
            "res3_down2": dict(initial_res_fac=3, max_scaling_down=2), "down4": dict(max_scaling_down=4)}
    for name, o in opts.items():
        cfg = synth.get_config("tiny_b", ctc={}, **o)
        w = {k: v.astype(np.float64) for k, v in synth.make_weights(cfg, bias_noise=0.1).items()}
        batch = synth.make_batch(cfg)
        depths = {"t%d" % d: d for d in range(1, cfg.L + 1)}
        states, lens, _ = om.encoder_fwd(w, batch["logmel"].astype(np.float64), batch["logmel_len"], depths, o)
        for d in range(1, cfg.L + 1):
            np.testing.assert_allclose(states[d], G["%s/states/%d" % (name, d)], rtol=0, atol=1e-13, err_msg="%s %d" % (name, d))
            assert np.array_equal(lens[d], G["%s/lens/%d" % (name, d)]), (name, d)


def test_dropout_placement_equals_the_executed_reference_training_graph():
    """out_prob = 0.8 / out_prob_dec = 0.7 (the reference trains with 0.9): its own get_cell / DropoutWrapper wiring
    executed with the oracle's Philox keep-masks injected on the encoder cells' and lm_cell's outputs and ZEROS on the
    decoder cell's output -- the reference never reads that output (attn_decoder.py:114-118 takes the cell state), so
    the oracle, which applies no mask there, must still reproduce states, logits and loss."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_dropout.npz"))
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    res = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks={}, want_grads=False,
                        out_prob=float(G["keep"][0]), out_prob_dec=float(G["keep"][1]), dropout_seed=int(G["seed"]))
    np.testing.assert_allclose(res["states"][cfg.L], G["states"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(res["logits"]["char"], G["logits"], rtol=0, atol=1e-12)
    assert abs(res["losses"]["char"] - float(G["loss"])) < 1e-12
    plain = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks={}, want_grads=False)
    assert abs(plain["losses"]["char"] - float(G["loss"])) > 1e-4            # the masks really acted


def test_stacked_decoder_dropout_equals_the_executed_reference_training_graph():
    """2-layer LSTM decoder with out_prob_dec = 0.7: every single cell of lm_cell and of the decoder cell is wrapped
    (decoder.py:53-68); masks injected per layer, zeros on the top decoder layer's never-read output."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_dropout_dec2.npz"))
    cfg = synth.get_config("tiny_dec2")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    res = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks={}, want_grads=False,
                        out_prob_dec=float(G["keep"]), dropout_seed=int(G["seed"]),
                        dec_params={"num_layers_dec": 2, "use_lstm": True})
    np.testing.assert_allclose(res["logits"]["char"], G["logits"], rtol=0, atol=1e-12)
    assert abs(res["losses"]["char"] - float(G["loss"])) < 1e-12


def test_reference_default_widths_equal_the_executed_reference_graph():
    """cfg-1 (the reference's base_params defaults: H = Hd = Hl = 256, A = 128, V = 1000, 4 pyramid layers): loss and
    sub-sampled logits / top-layer states of the executed reference graph."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "graph_cfg1.npz"))
    cfg = synth.get_config("cfg1")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    res = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks={}, want_grads=False)
    assert tuple(res["logits"]["char"].shape) == tuple(G["logits_shape"])
    np.testing.assert_allclose(res["logits"]["char"].reshape(-1)[::97], G["logits_sub"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(res["states"][cfg.L].reshape(-1)[::89], G["states_sub"], rtol=0, atol=1e-12)
    assert np.array_equal(res["lens"][cfg.L], G["lens"])
    assert abs(res["losses"]["char"] - float(G["loss"])) < 1e-11

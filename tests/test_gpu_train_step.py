"""Whole-step parity on the GPU: losses, logits and clipped gradients of one
teacher-forced fwd+bwd step against the CPU oracle (north-star: fp32 within 1e-4
relative)."""
import numpy as np
import pytest
import torch

from e2e_asr_b200 import ops, synth
from e2e_asr_b200.testing import build_model, compare_step
from oracle import model as om

pytestmark = pytest.mark.gpu
RTOL = 1e-4    # north-star tolerance for fp32 losses / logits / gradients


@pytest.mark.parametrize("cname,ctc", [("tiny", True), ("tiny_b", True), ("tiny", False), ("cfg1", True),
                                       ("wide_small", True)])
def test_train_step_matches_oracle(cname, ctc):
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0", ctc=ctc)
    model.run_step(batch)
    ops.check_device_errors("cuda:0")
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc if ctc else {})
    worst = compare_step(model, ref, rtol=RTOL)
    # reference-shaped public tensors
    d = cfg.L
    assert tuple(model.encoder_hidden_states[d].shape) == ref["states"][d].shape
    assert tuple(model.outputs["char"].shape) == ref["logits"]["char"].shape
    # second step on the same batch is bit-identical in the loss (buffers are re-zeroed)
    l1 = float(model.total_loss)
    model.run_step(batch)
    assert abs(float(model.total_loss) - l1) <= 1e-6 * abs(l1)
    print(cname, "worst grad rel err", worst)


def test_clipping_active_and_dense_norm_option():
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": 4}, ctc_tasks=cfg.ctc, max_gradient_norm=0.5)
    model = build_model(cfg, w, device="cuda:0")
    model.params.max_gradient_norm = 0.5
    model.run_step(batch)
    assert ref["norm"] > 0.5
    compare_step(model, ref, rtol=RTOL)
    model.params.tf_indexed_slices_norm = False
    model.run_step(batch)
    assert abs(float(model.grad_norm) - ref["dense_norm"]) < 1e-4 * ref["dense_norm"]


@pytest.mark.parametrize("mode,rtol", [("tf32x3", 1e-4), ("f16x2", 1e-4), ("bf16x2", 1e-4), ("bf16", 5e-2)])
def test_train_step_tensor_core_modes(mode, rtol):
    """cfg1 through the tcgen05 GEMMs.  tf32x3 is the fp32-accurate mode and must meet
    the same 1e-4 bar as FFMA, and so must bf16x2 (hi + lo bf16 operands, three products); bf16 is the reduced-precision mode: stated tolerance
    5e-2 on gradients/logits (relative to each tensor's max), 1e-2 on the loss."""
    cfg = synth.get_config("cfg1")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    ops.set_gemm_mode(mode)
    try:
        model = build_model(cfg, w, device="cuda:0")
        model.run_step(batch)
        ops.check_device_errors("cuda:0")
        if mode == "bf16":
            for t, l in ref["losses"].items():
                assert abs(float(model.losses[t]) - l) <= 1e-2 * max(1.0, abs(l))
            grads = model.gradients()
            gmax = max(float(np.abs(g).max()) for g in ref["clipped"].values())
            for k, g in ref["clipped"].items():
                err = float(np.abs(grads[k].astype(np.float64) - g).max())
                assert err <= rtol * max(float(np.abs(g).max()), 1e-2 * gmax), (k, err)
        else:
            worst = compare_step(model, ref, rtol=rtol)
            print(mode, "worst grad rel err", worst)
    finally:
        ops.set_gemm_mode("fp32")


@pytest.mark.parametrize("cname", ["tiny", "tiny_b", "cfg1"])
def test_greedy_decode_ids_bit_exact(cname):
    """Eval-mode graph: always max_output steps, argmax feedback.  Token ids
    extracted as eval_model.py:84-87 must equal the oracle's greedy ids exactly."""
    from oracle import beam as ob
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] = w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] * 6.0
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0", isTraining=False, ctc=False)
    model.run_step(batch)
    logits = model.outputs["char"].detach().cpu().numpy()
    U = cfg.U
    assert logits.shape == (U * cfg.B, cfg.V)
    ids = ob.greedy_ids_from_logits(logits, cfg.B)
    # oracle: same encoder states (float64), greedy decoder with len := max_output
    W64 = {k: v.astype(np.float64) for k, v in w.items()}
    states, lens_d, _ = om.encoder_fwd(W64, batch["logmel"].astype(np.float64), batch["logmel_len"], {"char": cfg.L})
    ref_logits, _ = om.attn_decoder_fwd(W64, "char", batch["char"].T, np.full(cfg.B, U), states[cfg.L],
                                        lens_d[cfg.L], mode="greedy", max_steps=U)
    ref_ids = ob.greedy_ids_from_logits(ref_logits, cfg.B)
    np.testing.assert_array_equal(ids, ref_ids)
    e = np.abs(logits - ref_logits).max() / np.abs(ref_logits).max()
    assert e < 1e-4


@pytest.mark.parametrize("impl", ["persist", "loop"])
@pytest.mark.parametrize("cname", ["tiny_b", "cfg1"])
def test_train_step_with_dropout_matches_oracle(cname, impl):
    """out_prob / out_prob_dec < 1 (the reference's training defaults are 0.9): the Philox masks are a builder-defined
    stateless function of (seed, stream, flat index), restated in the oracle, so the dropped step must meet the same
    1e-4 bar.  The second step uses another key (global_step is mixed in)."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0")
    model.params.encoder_params.out_prob = 0.8
    model.params.decoder_params["char"].out_prob_dec = 0.7
    model.params.dropout_seed = 11
    ops.set_decoder_impl(impl)            # the persistent decoder kernels and the per-step fallback (D = 1024 shapes)
    try:
        for step in range(2):
            model.run_step(batch)
            ops.check_device_errors("cuda:0")
            ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, out_prob=0.8,
                                out_prob_dec=0.7, dropout_seed=11 * 1000003 + step)
            compare_step(model, ref, rtol=RTOL)
    finally:
        ops.set_decoder_impl("persist")
    nodrop = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    assert abs(ref["total_loss"] - nodrop["total_loss"]) > 1e-3


@pytest.mark.parametrize("cname", ["tiny_b", "wide_small"])
def test_graphed_step_with_the_per_step_decoder_kernels(cname):
    """The capture must also work when the decoder runs the per-step kernels (shapes the persistent loop cannot hold,
    e.g. cfg-5's D = 1024): that path never touches the "dec" side stream, which the end-of-step join then has to
    leave alone (waiting on an idle, uncaptured stream aborts a capture)."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    ops.set_decoder_impl("loop")
    try:
        model = build_model(cfg, w, device="cuda:0")
        gs = model.graphed_step(batch)
        gs.step(batch)
        ops.check_device_errors("cuda:0")
        compare_step(model, ref, rtol=RTOL)
    finally:
        ops.set_decoder_impl("persist")


@pytest.mark.parametrize("cname", ["tiny_b", "cfg1"])
def test_graphed_step_with_dropout_draws_a_new_mask_every_replay(cname):
    """The reference's training defaults keep 0.9 of the outputs: the step captured WITH dropout must draw the mask
    of ITS step at every replay -- the kernels read the Philox key from a device word rewritten before each launch --
    and each replay must meet the 1e-4 bar against the oracle run with that step's key."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0")
    model.params.encoder_params.out_prob = 0.9
    model.params.decoder_params["char"].out_prob_dec = 0.9
    model.params.dropout_seed = 7
    gs = model.graphed_step(batch)
    assert model.global_step == 0
    losses = []
    for step in range(3):
        gs.step(batch)
        ops.check_device_errors("cuda:0")
        ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, out_prob=0.9, out_prob_dec=0.9,
                            dropout_seed=7 * 1000003 + step)
        compare_step(model, ref, rtol=RTOL)
        losses.append(ref["total_loss"])
    assert len(set(round(l, 5) for l in losses)) == 3          # three different masks


@pytest.mark.parametrize("cname,keep", [("tiny_b", 1.0), ("tiny_b", 0.7), ("cfg1", 0.9)])
def test_scheduled_sampling_matches_oracle(cname, keep):
    """samp_prob > 0 (attn_decoder.py:130-139, reference default 0.1): the per-step scalar draw and the multinomial
    draw are builder-defined Philox functions restated in the oracle, so the realised input ids must be IDENTICAL
    and the step on them must meet the 1e-4 bar (with and without LM-output dropout in the sampling pass)."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0")
    model.params.decoder_params["char"].samp_prob = 0.5
    model.params.decoder_params["char"].out_prob_dec = keep
    model.params.dropout_seed = 5
    sampled_any = False
    for step in range(3):
        model.run_step(batch)
        ops.check_device_errors("cuda:0")
        kw = dict(num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, out_prob_dec=keep, dropout_seed=5 * 1000003 + step)
        ref = om.train_step(w, batch, samp_prob=0.5, **kw)
        ids = model.decoder["char"].stash["realized_ids"].cpu().numpy()
        ids_equal_up_to_bin_edges(ids, ref["realized_ids"]["char"])
        teacher = np.asarray(batch["char"]).T[:ids.shape[0]]
        sampled_any |= bool((ids != teacher).any())
        if not np.array_equal(ids, ref["realized_ids"]["char"]):
            ref = om.train_step(w, batch, forced_inputs={"char": ids}, **kw)      # the step on the ids realised here
        compare_step(model, ref, rtol=RTOL)
    assert sampled_any          # the rule really replaced ground-truth inputs


def ids_equal_up_to_bin_edges(ids, ref_ids):
    """The multinomial draw is an inverse-CDF lookup (first index whose cumulative mass exceeds u * total): logits that
    differ in the 6th digit move a draw that falls within that distance of a bin edge to the ADJACENT id (with V = 1000
    near-uniform classes about one draw in a hundred).  From there on that utterance is fed another embedding and its
    later draws are no longer comparable.  So: per utterance the realised ids must be identical up to the first
    difference, and that first difference must be an adjacent id.  (The step itself is then checked against the oracle
    run on the ids realised here.)"""
    assert ids.shape == ref_ids.shape
    for b in range(ids.shape[1]):
        d = np.flatnonzero(ids[:, b] != ref_ids[:, b])
        if len(d):
            t = int(d[0])
            assert abs(int(ids[t, b]) - int(ref_ids[t, b])) == 1, (b, t, ids[t, b], ref_ids[t, b])


def test_sample_rows_kernel_matches_oracle():
    """e2e_sample_rows against the oracle's inverse-CDF draw, plus the empirical distribution of many draws."""
    from e2e_asr_b200._lib import call
    g = torch.Generator().manual_seed(0)
    B, V = 64, 37
    lg = (torch.randn(B, V, generator=g) * 2.0).cuda()
    out = torch.empty(B, dtype=torch.int64, device="cuda")
    call("e2e_sample_rows", B, V, lg, V, 1234567891011, 301, 640, out)
    assert np.array_equal(out.cpu().numpy(), om.sample_rows(lg.cpu().numpy(), 1234567891011, 301, 640))
    # distribution: 20000 rows of the same logits
    n = 20000
    one = torch.tensor([0.0, 1.0, 2.0, -1.0, 0.5]).cuda()
    rows = one.repeat(n, 1).contiguous()
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    call("e2e_sample_rows", n, 5, rows, 5, 7, 300, 0, out)
    freq = np.bincount(out.cpu().numpy(), minlength=5) / n
    p = torch.softmax(one.cpu().double(), 0).numpy()
    assert np.abs(freq - p).max() < 0.015


@pytest.mark.parametrize("cname", ["tiny_b", "cfg1"])
def test_graphed_step_matches_eager_and_oracle(cname):
    """GraphedStep: the step captured in a CUDA graph (side streams included) and replayed (a) on the captured batch
    and (b) on another batch of the same shape refilled into the static input buffers must reproduce the eager step,
    which is itself checked against the oracle.  A batch with other maximum lengths must be refused."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    perm = np.roll(np.arange(cfg.B), 1)
    batch2 = {k: (np.asarray(v)[perm].copy() if hasattr(v, "shape") and np.asarray(v).shape[:1] == (cfg.B,) else v)
              for k, v in batch.items()}
    batch2["logmel"] = (batch2["logmel"] * 0.9 + 0.05).astype(batch["logmel"].dtype)
    model = build_model(cfg, w, device="cuda:0")
    gs = model.graphed_step(batch)
    assert gs.launches_per_step > 20
    for b in (batch, batch2, batch):
        gs.step(b)
        ops.check_device_errors("cuda:0")
        ref = om.train_step(w, b, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
        compare_step(model, ref, rtol=RTOL)
    # double-buffered input pipeline: prefetch() stages the next batch on a copy stream, step() consumes it
    ref2 = om.train_step(w, batch2, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    gs.prefetch(batch2)
    gs.step()
    gs.prefetch(batch)
    compare_step(model, ref2, rtol=RTOL)
    gs.step()
    compare_step(model, ref, rtol=RTOL)
    eager = build_model(cfg, w, device="cuda:0")
    eager.run_step(batch)
    # eager vs replay: the same kernels; split-K sums use atomics, so not bit-identical (same floor as compare_step)
    g0, g1 = eager.gradients(), model.gradients()
    gmax = max(float(np.abs(g).max()) for g in g0.values())
    for k in g0:
        assert np.abs(g0[k] - g1[k]).max() <= 2e-5 * max(np.abs(g0[k]).max(), 1e-4 * gmax), k
    short = {k: v for k, v in batch.items()}
    short["logmel_len"] = np.maximum(np.asarray(batch["logmel_len"]) - 1, 1)
    with pytest.raises(ValueError):
        gs.step(short)


@pytest.mark.parametrize("cname,mode", [("tiny_uni", "fp32"), ("uni256", "fp32"), ("uni256", "tf32x3")])
def test_unidirectional_encoder_matches_oracle(cname, mode):
    """bi_dir=False (encoder.py:86-89): forward-only tf.nn.dynamic_rnn layers, D = H, variables named
    RNNLayer<d>/<d>/basic_lstm_cell/*; the recurrence kernels run with one direction (H=16: generic kernel, H=256:
    the cluster kernel)."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, enc_params={"bi_dir": False})
    ops.set_gemm_mode(mode)
    try:
        model = build_model(cfg, w, device="cuda:0")
        assert not model.params.encoder_params.bi_dir
        for _ in range(2):
            model.run_step(batch)
            ops.check_device_errors("cuda:0")
            compare_step(model, ref, rtol=RTOL)
        assert model.encoder_hidden_states[cfg.L].shape[2] == cfg.H
    finally:
        ops.set_gemm_mode("fp32")


@pytest.mark.parametrize("cname,mode", [("tiny_gru", "fp32"), ("gru256", "fp32"), ("gru256", "tf32x3")])
def test_gru_encoder_matches_oracle(cname, mode):
    """use_lstm=False in the encoder (encoder.py:27,48: tf.nn.rnn_cell.GRUCell -- the reference's class_params
    default): gates / candidate kernels with TF's variable names, recurrence in e2e_gru_rec_fwd/bwd, against the
    oracle's GRU restatement (itself checked by finite differences)."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, enc_params={"use_lstm": False})
    ops.set_gemm_mode(mode)
    try:
        model = build_model(cfg, w, device="cuda:0")
        assert not model.params.encoder_params.use_lstm
        for _ in range(2):
            model.run_step(batch)
            ops.check_device_errors("cuda:0")
            compare_step(model, ref, rtol=RTOL)
    finally:
        ops.set_gemm_mode("fp32")


@pytest.mark.parametrize("cname,ctc", [("tiny_b", True), ("cfg1", False)])
def test_multitask_char_and_phone_decoders_match_oracle(cname, ctc):
    """The reference's multitask setup (main.py:89-93): a character attention decoder on the top encoder layer and a
    PHONE attention decoder on a lower one (own vocabulary, own variables under rnn_decoder_phone), losses averaged
    over tasks (seq2seq_model.py:140-144) -- eager and replayed from the CUDA graph."""
    cfg = synth.get_config(cname, V_phone=13 if cname == "tiny_b" else 48)
    tasks = ("char", "phone")
    depth = {"char": cfg.L, "phone": cfg.L - 1}
    w = synth.make_weights(cfg, tasks=tasks, bias_noise=0.1)
    batch = synth.make_batch(cfg, tasks=tasks)
    ref = om.train_step(w, batch, tasks=tasks, num_layers=depth, ctc_tasks=cfg.ctc if ctc else {})
    assert abs(ref["losses"]["char"] - ref["losses"]["phone"]) > 1e-3
    model = build_model(cfg, w, device="cuda:0", tasks=tasks, num_layers=depth, ctc=ctc)
    for _ in range(2):
        model.run_step(batch)
        ops.check_device_errors("cuda:0")
        compare_step(model, ref, rtol=RTOL)
    gs = model.graphed_step(batch)
    gs.step(batch)
    compare_step(model, ref, rtol=RTOL)


@pytest.mark.parametrize("cname,keep", [("tiny_dec2", 1.0), ("tiny_decgru", 1.0), ("tiny_decgru2", 1.0),
                                        ("tiny_dec2", 0.7), ("tiny_decgru2", 0.6)])
def test_general_decoder_cells_match_oracle(cname, keep):
    """decoder.py:49-82: lm_cell and the decoder cell as MultiRNNCell stacks (num_layers_dec = 2) and / or GRU cells
    (use_lstm=False); the attention query is the last layer's c (LSTM) or state (GRU).  Step-by-step path
    (ops.attn_decoder_stepwise) against the oracle's general restatement, which is bit-identical to the pinned one
    for the single LSTM cell and checked by finite differences otherwise."""
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    dec = {"num_layers_dec": cfg.get("dec_layers", 1), "use_lstm": cfg.get("dec_lstm", True)}
    model = build_model(cfg, w, device="cuda:0")
    assert model.decoder["char"].general_cells()
    model.params.decoder_params["char"].out_prob_dec = keep      # DropoutWrapper on every single cell's output
    model.params.dropout_seed = 3
    for step in range(2):
        ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, dec_params=dec,
                            out_prob_dec=keep, dropout_seed=3 * 1000003 + step)
        model.run_step(batch)
        ops.check_device_errors("cuda:0")
        compare_step(model, ref, rtol=RTOL)


@pytest.mark.parametrize("cname,keep", [("tiny_dec2", 1.0), ("tiny_decgru", 1.0), ("tiny_decgru2", 0.7)])
def test_general_decoder_cells_scheduled_sampling_and_greedy(cname, keep):
    """The decoder's other input rules for stacked / GRU cells (decoder.py:139-180, attn_decoder.py:127-139): scheduled
    sampling (realised input ids IDENTICAL to the oracle's, step within 1e-4) and eval-mode greedy decoding (ids
    bit-exact); the oracle's general restatement of both is pinned to the executed reference (graph_modes_general)."""
    from oracle import beam as ob
    cfg = synth.get_config(cname)
    w = synth.make_weights(cfg, bias_noise=0.1)
    w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] = w["model/rnn_decoder_char/rnn/OutputProjection/kernel"] * 6.0
    batch = synth.make_batch(cfg)
    dec = {"num_layers_dec": cfg.get("dec_layers", 1), "use_lstm": cfg.get("dec_lstm", True)}
    model = build_model(cfg, w, device="cuda:0")
    model.params.decoder_params["char"].samp_prob = 0.5
    model.params.decoder_params["char"].out_prob_dec = keep
    model.params.dropout_seed = 9
    sampled_any = False
    for step in range(3):
        model.run_step(batch)
        ops.check_device_errors("cuda:0")
        ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, dec_params=dec, out_prob_dec=keep,
                            dropout_seed=9 * 1000003 + step, samp_prob=0.5)
        ids = model.decoder["char"].stash["realized_ids"].cpu().numpy()
        assert np.array_equal(ids, ref["realized_ids"]["char"])
        sampled_any |= bool((ids != np.asarray(batch["char"]).T[:ids.shape[0]]).any())
        compare_step(model, ref, rtol=RTOL)
    assert sampled_any
    # eval mode
    ev = build_model(cfg, w, device="cuda:0", isTraining=False, ctc=False)
    ev.run_step(batch)
    logits = ev.outputs["char"].detach().cpu().numpy()
    assert logits.shape == (cfg.U * cfg.B, cfg.V)
    W64 = {k: v.astype(np.float64) for k, v in w.items()}
    states, lens_d, _ = om.encoder_fwd(W64, batch["logmel"].astype(np.float64), batch["logmel_len"], {"char": cfg.L})
    ref_logits, _ = om.attn_decoder_general(W64, "char", batch["char"].T, np.full(cfg.B, cfg.U), states[cfg.L],
                                            lens_d[cfg.L], dec["num_layers_dec"], dec["use_lstm"], mode="greedy",
                                            max_steps=cfg.U)
    np.testing.assert_array_equal(ob.greedy_ids_from_logits(logits, cfg.B), ob.greedy_ids_from_logits(ref_logits, cfg.B))
    assert np.abs(logits - ref_logits).max() / np.abs(ref_logits).max() < 1e-4


@pytest.mark.parametrize("graphed", [False, True])
def test_adam_updates_match_oracle(graphed):
    """apply_updates=True: three Adam steps (fused flat-buffer kernel) against the oracle's TF-formula Adam, eager and
    with the step replayed from its CUDA graph (capturing must not touch the parameters or the optimiser state; the
    replays read the updated weights in place)."""
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0")
    model.params.apply_updates = True
    gs = model.graphed_step(batch) if graphed else None
    if graphed:
        assert model.global_step == 0 and model._adam is None
    wref = {k: v.astype(np.float64) for k, v in w.items()}
    state = {}
    for step in range(3):
        if graphed:
            gs.step(batch)
        else:
            model.run_step(batch)
        ref = om.train_step(wref, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
        wref = om.adam_step(wref, ref["clipped"], state, 1e-3)
    got = model.variables.state_dict()
    for k, v in wref.items():
        err = float(np.abs(got[k] - v).max())
        # a step moves a weight by ~lr = 1e-3: 1e-6 absolute = 1e-3 of the update
        assert err < 2e-6, (k, err)
        assert float(np.abs(got[k] - w[k]).max()) > 1e-4      # the parameters really moved


def test_training_loop_overfits_one_batch_like_the_oracle():
    """60 Adam steps (lr 1e-2) on one batch through graph replays: the loss must fall from ~7.6 to ~0.2 and follow the
    float64 oracle's own trajectory (fp32 rounding is amplified along 60 updates, hence the loose 5 % band)."""
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg)
    batch = synth.make_batch(cfg)
    model = build_model(cfg, w, device="cuda:0")
    model.params.apply_updates = True
    model.learning_rate = 1e-2
    gs = model.graphed_step(batch)
    wref = {k: v.astype(np.float64) for k, v in w.items()}
    state = {}
    for step in range(60):
        gs.step(batch)
        ref = om.train_step(wref, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
        wref = om.adam_step(wref, ref["clipped"], state, 1e-2)
        if step in (0, 20, 40, 59):
            got = float(model.total_loss)
            assert abs(got - ref["total_loss"]) <= 0.05 * max(1.0, ref["total_loss"]), (step, got, ref["total_loss"])
    assert float(model.total_loss) < 0.5
    ops.check_device_errors("cuda:0")

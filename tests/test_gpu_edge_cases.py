"""Edge cases of the training step against the oracle: single-utterance batches, one-frame utterances (every pyramid
level has length 1), one-token targets (EOS only), ragged lengths with odd maxima, and the encoder's input options
(frame stacking seq2seq_model.py:164-183, initial stride encoder.py:149-153, no pyramid skip_step=1)."""
import numpy as np
import pytest

from e2e_asr_b200 import ops, synth
from e2e_asr_b200.data_utils import EOS_ID, GO_ID, PAD_ID
from e2e_asr_b200.testing import build_model, compare_step
from oracle import model as om

pytestmark = pytest.mark.gpu


def crafted_batch(cfg, frame_lens, target_lens, seed=0):
    """A batch with exactly these lengths (logmel zero past each length, ids GO .. EOS PAD..)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    B, T, U = len(frame_lens), max(frame_lens), max(target_lens)
    logmel = rng.standard_normal((B, T, cfg.F)).astype(np.float32)
    ids = np.full((B, U + 1), PAD_ID, np.int64)
    for b in range(B):
        logmel[b, frame_lens[b]:] = 0.0
        n = target_lens[b]
        ids[b, 0] = GO_ID
        ids[b, 1:n] = rng.integers(3, cfg.V, size=n - 1)
        ids[b, n] = EOS_ID
    batch = {"logmel": logmel, "logmel_len": np.asarray(frame_lens, np.int64), "char": ids,
             "char_len": np.asarray(target_lens, np.int64), "utt_id": np.array(["u%d" % b for b in range(B)])}
    for task, (depth, vocab) in cfg.ctc.items():
        dl = synth.pyramid_lens(frame_lens, synth.depth_reductions(cfg, depth))
        ll = np.maximum(1, np.minimum(dl // 2, 3)).astype(np.int64)       # feasible: label length <= frames
        lab = np.zeros((B, int(ll.max())), np.int64)
        for b in range(B):
            lab[b, :ll[b]] = rng.integers(0, vocab, size=int(ll[b]))
        batch[task], batch[task + "_len"] = lab, ll
    return batch


@pytest.mark.parametrize("frame_lens,target_lens", [
    ([1], [1]),                          # one utterance, one frame, EOS only
    ([9], [4]),                          # one utterance
    ([1, 2, 23], [1, 6, 2]),             # one-frame and two-frame utterances next to a long one; odd maximum
    ([17, 17, 17, 17], [3, 3, 3, 3]),    # no padding at all
    ([5, 31, 8, 31, 2], [6, 1, 6, 2, 5]),
])
def test_ragged_and_degenerate_lengths(frame_lens, target_lens):
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = crafted_batch(cfg, frame_lens, target_lens)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    assert np.isfinite(ref["total_loss"])
    model = build_model(cfg, w, device="cuda:0")
    for _ in range(2):
        model.run_step(batch)
        ops.check_device_errors("cuda:0")
        compare_step(model, ref, rtol=1e-4)
    # the same step replayed from its CUDA graph
    gs = model.graphed_step(batch)
    gs.step(batch)
    compare_step(model, ref, rtol=1e-4)


@pytest.mark.parametrize("enc", [dict(stack_cons=3), dict(initial_res_fac=2), dict(skip_step=1),
                                 dict(stack_cons=2, initial_res_fac=3, max_scaling_down=2)])
def test_encoder_input_options(enc):
    """stack_cons (frame stacking), initial_res_fac (input stride), skip_step=1 (no pyramid), max_scaling_down."""
    base = synth.get_config("tiny_b")
    stack = enc.get("stack_cons", 1)
    # weights for the architecture these options produce: stacked feature width, fewer / no pyramid steps
    cfg = synth.get_config("tiny_b", F=base.F * stack, ctc={}, **{k: v for k, v in enc.items() if k != "stack_cons"})
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = crafted_batch(base, [37, 20, 9, 33, 1], [5, 9, 2, 1, 4])
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks={}, enc_params=enc)
    model = build_model(cfg, w, device="cuda:0", ctc=False)
    for k, v in enc.items():
        model.params.encoder_params[k] = v
    model.run_step(batch)
    ops.check_device_errors("cuda:0")
    compare_step(model, ref, rtol=1e-4)


@pytest.mark.parametrize("frame_lens,target_lens", [
    ([57], [9]),                                             # one utterance through the flagship kernels
    ([40, 3, 64, 17, 1, 64, 33, 8, 21, 5, 64, 2, 47, 30, 11, 64, 9], [12, 1, 3, 12, 2, 7, 1, 12, 5, 9, 4, 12, 6, 2, 8, 3, 10]),
])
@pytest.mark.parametrize("mode", ["fp32", "tf32x3"])
def test_ragged_batches_at_reference_widths(frame_lens, target_lens, mode):
    """The reference's default widths (H = Hd = Hl = 256, A = 128, V = 1000: the cluster-resident recurrence, the
    persistent decoder loop, tcgen05 GEMMs) on a single utterance and on 17 utterances (not a multiple of the 16-row
    batch slices) with lengths from 1 frame / 1 token to the maximum."""
    cfg = synth.get_config("cfg1")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = crafted_batch(cfg, frame_lens, target_lens)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    ops.set_gemm_mode(mode)
    try:
        model = build_model(cfg, w, device="cuda:0")
        for _ in range(2):
            model.run_step(batch)
            ops.check_device_errors("cuda:0")
            compare_step(model, ref, rtol=1e-4)
        gs = model.graphed_step(batch)
        gs.step(batch)
        compare_step(model, ref, rtol=1e-4)
    finally:
        ops.set_gemm_mode("fp32")


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
def test_distinct_batches_back_to_back_from_host(graph):
    """The host -> device staging is asynchronous (pinned buffer -> non_blocking copy) and the host runs one step ahead:
    step N must train on batch N even though batch N+1 is staged while it runs.  Different batches of the same shape
    are fed back to back WITHOUT any synchronisation in between; every step's loss must equal the loss of that batch
    fed from device-resident (`prepared=`) inputs."""
    import torch
    cfg = synth.get_config("cfg1", B=8, T=120)
    w = synth.make_weights(cfg)
    base = synth.make_batch(cfg, seed=100)
    batches = []
    for i in range(4):          # same shapes, lengths and maxima; other features and other (valid) token ids
        b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in base.items()}
        b["logmel"] = (base["logmel"] * (1.0 + 0.25 * i)).astype(np.float32)
        tok = b["char"] >= 3
        b["char"][tok] = (b["char"][tok] - 3 + 7 * i) % (cfg.V - 3) + 3
        batches.append(b)
    ops.set_gemm_mode("tf32x3")
    try:
        model = build_model(cfg, w, device="cuda:0")
        expect = []
        for b in batches:
            model.run_step(b)
            torch.cuda.synchronize()
            expect.append(float(model.total_loss))
        assert len(set(round(e, 6) for e in expect)) == len(expect), "the batches must differ"
        step = model.graphed_step(batches[0]) if graph else model.run_step
        got = []
        for rep in range(3):
            for b in batches:
                step(b)
                got.append(model.total_loss.clone())  # device scalars (a replayed graph rewrites ONE tensor): no host
                                                      # synchronisation between steps
        torch.cuda.synchronize()
        got = [float(g) for g in got]
        for i, g in enumerate(got):
            assert abs(g - expect[i % 4]) <= 1e-6 * abs(expect[i % 4]), (i, g, expect[i % 4])
    finally:
        ops.set_gemm_mode("fp32")


def test_one_captured_step_serves_a_length_bucket():
    """The reference's training loop draws batches from length buckets (train.py:44,108-119): every batch of a bucket
    has its own maximum lengths.  GraphedStep(bucket=True) derives the loop bounds of the recurrences, the decoder loop
    and the CTC state space from the PADDED shapes (the extra steps are masked), so ONE captured step is replayed on
    batches with different maxima, each padded to the bucket's shapes -- and each must meet the oracle's result for the
    unpadded batch."""
    import torch
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    cap = crafted_batch(cfg, [37, 30, 21, 37, 12], [9, 4, 9, 2, 6], seed=1)          # the bucket's bounds
    batches = [crafted_batch(cfg, [29, 30, 21, 8, 12], [5, 4, 7, 2, 6], seed=2),      # shorter maxima (odd)
               crafted_batch(cfg, [4, 3, 2, 1, 5], [1, 2, 1, 1, 3], seed=3),          # far below the bounds
               cap]
    model = build_model(cfg, w, device="cuda:0")
    step = model.graphed_step(cap, bucket=True)
    for b in batches * 2:
        ref = om.train_step(w, b, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
        step(b)
        torch.cuda.synchronize()
        ops.check_device_errors("cuda:0")
        compare_step(model, ref, rtol=1e-4, check_logits=False)
        U = int(b["char_len"].max())
        got = model.outputs["char"].detach().cpu().numpy().reshape(-1, len(b["char_len"]), cfg.V)
        assert np.abs(got[:U].reshape(-1, cfg.V) - ref["logits"]["char"]).max() <= 1e-4 * np.abs(ref["logits"]["char"]).max()
        assert not got[U:].any()                       # padded steps emit zeros (finished rows, raw_rnn)
    too_long = crafted_batch(cfg, [38, 30, 21, 37, 12], [9, 4, 9, 2, 6], seed=4)
    with pytest.raises(ValueError, match="does not fit"):
        step(too_long)


def test_ctc_head_with_a_task_name_the_encoder_does_not_keep_time_major():
    """The reference's encoder keeps the time-major view only for tasks named "state" (encoder.py:143-144; here also
    "*_ctc"); an auxiliary CTC head under any other task name reads its layer from the batch-major attention states."""
    cfg = synth.get_config("tiny_b", ctc={"phone": (3, 6), "state": (2, 9)})
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    ref = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    model = build_model(cfg, w, device="cuda:0")
    model.run_step(batch)
    ops.check_device_errors("cuda:0")
    compare_step(model, ref, rtol=1e-4)

"""Scheduled sampling (attn_decoder.py:130-139, decoder.py:155-180), CPU side: the oracle's restatement and the
product's host-side scalar Philox draw."""
import numpy as np

from e2e_asr_b200 import synth
from e2e_asr_b200.host_utils import philox_uniform
from oracle import model as om


def test_host_philox_matches_oracle_philox():
    for counter, offset, seed in [(0, 0, 0), (1, 200, 5), (77, 203, 11 * 1000003 + 2), (2 ** 31 + 5, 9, 2 ** 40 + 3)]:
        w = om.philox4x32_10(np.array([counter], np.uint64), np.array([offset], np.uint64), np.zeros(1, np.uint64),
                             np.zeros(1, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0]
        assert philox_uniform(counter, offset, seed) == float(w[0]) * 2.0 ** -32


def test_sample_decisions_rate_and_first_step():
    use = om.sample_decisions(20000, 0.1, 3, 0)
    assert not use[0]                       # step 0 always reads GO
    assert abs(use.mean() - 0.1) < 0.01
    assert not om.sample_decisions(50, 0.0, 3, 0).any()
    assert om.sample_decisions(50, 1.0, 3, 0)[1:].all()


def test_sample_rows_distribution():
    lg = np.tile(np.array([0.0, 1.0, 2.0, -1.0, 0.5]), (20000, 1))
    ids = om.sample_rows(lg, 7, 300, 0)
    p = np.exp(lg[0]) / np.exp(lg[0]).sum()
    assert np.abs(np.bincount(ids, minlength=5) / len(ids) - p).max() < 0.015
    # a one-hot posterior always returns its mode
    sharp = np.full((8, 6), -1e4)
    sharp[np.arange(8), np.arange(8) % 6] = 0.0
    assert np.array_equal(om.sample_rows(sharp, 1, 300, 0), np.arange(8) % 6)


def test_sampled_step_is_the_teacher_step_on_the_realised_ids():
    """The draw carries no gradient: a sampled step equals the teacher-forced step fed the realised ids."""
    cfg = synth.get_config("tiny_b")
    w = synth.make_weights(cfg, bias_noise=0.1)
    batch = synth.make_batch(cfg)
    a = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc, samp_prob=0.6, dropout_seed=4)
    ids = a["realized_ids"]["char"]                       # [U, B]
    teacher = np.asarray(batch["char"]).T
    assert (ids != teacher[:ids.shape[0]]).any()
    assert np.array_equal(ids[0], teacher[0])
    # same step with the realised ids as inputs and the ORIGINAL targets
    W = {k: np.asarray(v, np.float64) for k, v in w.items()}
    x = om.stack_frames(np.asarray(batch["logmel"], np.float64), 1)
    lens = np.asarray(batch["logmel_len"], np.int64)
    states, lens_d, _ = om.encoder_fwd(W, x, lens, {"char": cfg.L}, None)
    seq_len = np.asarray(batch["char_len"], np.int64)
    fed = teacher.copy()
    fed[:ids.shape[0]] = ids
    lg, _ = om.attn_decoder_fwd(W, "char", fed, seq_len, states[cfg.L], lens_d[cfg.L])
    loss, _ = om.cross_entropy_loss(lg, teacher[1:int(seq_len.max()) + 1], seq_len)
    assert abs(loss - a["losses"]["char"]) < 1e-12
    # and samp_prob = 0 is the plain teacher-forced step
    b = om.train_step(w, batch, num_layers={"char": cfg.L}, ctc_tasks=cfg.ctc)
    assert np.array_equal(b["realized_ids"]["char"], teacher[:ids.shape[0]])
    assert abs(b["total_loss"] - a["total_loss"]) > 1e-6

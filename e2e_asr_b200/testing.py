"""Helpers shared by tests, smoke() and bench.py: build a Seq2SeqModel for a
synthetic config and compare one step against an oracle result dict.  (The
oracle itself is NOT imported here; callers pass its output in.)"""
import copy

import numpy as np
import torch

from . import synth
from .seq2seq_model import Seq2SeqModel
from .variables import VariableStore


def model_params(cfg, tasks=("char",), ctc=True, avg=True, num_layers=None):
    p = Seq2SeqModel.class_params()
    p.tasks = list(tasks)
    p.num_layers = {t: (num_layers or {}).get(t, cfg.L) for t in tasks}
    p.max_output = {t: cfg.U for t in tasks}
    p.avg = avg
    ep = p.encoder_params
    ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, bool(cfg.get("enc_lstm", True)), 1.0
    ep.bi_dir = bool(cfg.get("bi_dir", True))
    dp = {}
    for t in tasks:
        d = copy.deepcopy(p.decoder_params["char"])
        d.out_prob_dec, d.samp_prob = 1.0, 0.0
        d.hidden_size_dec, d.emb_size = cfg.Hd, cfg.E
        d.vocab_size = cfg.V if t == "char" else cfg.get("V_" + t, cfg.V)
        d.attention_vec_size, d.lm_hidden_size, d.max_output = cfg.A, cfg.Hl, cfg.U
        d.num_layers_dec, d.use_lstm = int(cfg.get("dec_layers", 1)), bool(cfg.get("dec_lstm", True))
        dp[t] = d
    p.decoder_params = dp
    p.ctc_tasks = {}
    if ctc:
        for t, (depth, vocab) in cfg.ctc.items():
            p.ctc_tasks[t] = vocab
            p.num_layers[t] = depth
    return p


def build_model(cfg, weights=None, device="cuda", isTraining=True, ctc=True, reducer=None, capacity=None,
                tasks=("char",), num_layers=None):
    if capacity is None:
        n = sum(int(np.prod(v.shape)) + 4 for v in (weights or synth.make_weights(cfg)).values())
        capacity = n + 1024
    vs = VariableStore(device, capacity=capacity)
    if weights is not None:
        vs.load(weights)
    return Seq2SeqModel(None, isTraining=isTraining,
                        params=model_params(cfg, tasks=tasks, ctc=ctc, num_layers=num_layers), variables=vs,
                        device=device, reducer=reducer)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def compare_step(model, ref, rtol=1e-4, check_logits=True):
    """Asserts losses / logits / clipped grads of `model` (after run_step) match the
    oracle dict `ref` within `rtol` (max-abs error relative to the tensor's max-abs,
    the north-star's "1e-4 relative").  Returns the worst gradient error."""
    for t, l in ref["losses"].items():
        got = float(model.losses[t].detach())
        assert abs(got - l) <= rtol * max(1.0, abs(l)), ("loss", t, got, l)
    assert abs(float(model.total_loss) - ref["total_loss"]) <= rtol * max(1.0, abs(ref["total_loss"]))
    if check_logits:
        for t in model.params.tasks:
            e = rel_err(model.outputs[t].detach().cpu().numpy(), ref["logits"][t])
            assert e <= rtol, ("logits", t, e)
    norm = float(model.grad_norm)
    assert abs(norm - ref["norm"]) <= rtol * max(1.0, ref["norm"]), ("norm", norm, ref["norm"])
    worst = 0.0
    grads = model.gradients()
    gmax = max(float(np.abs(g).max()) for g in ref["clipped"].values())
    for k, g in ref["clipped"].items():
        # per-variable max-abs error relative to that variable's max-abs value; variables
        # whose whole gradient is below 1e-4 of the largest gradient entry of the model
        # (pure cancellation residue -- the float32 oracle itself is only ~2e-4 accurate
        # on them) are measured against that floor instead.
        err = float(np.abs(np.asarray(grads[k], np.float64) - g).max())
        e = err / max(float(np.abs(g).max()), 1e-4 * gmax)
        worst = max(worst, e)
        assert e <= rtol, ("grad", k, e)
    return worst

"""tests/golden/grad_*.npz: gradients of the reference's OWN forward graph.  encoder.py / decoder.py / attn_decoder.py /
losses.py are executed unmodified on the torch-backed TensorFlow stand-in (torch_tf.py) and torch.autograd
differentiates the resulting loss with respect to every weight -- what tf.gradients(total_loss, trainable_vars) does in
seq2seq_model.py:148.  These pin the oracle's hand-derived backward pass (tests/test_graph_golden_cpu.py).
Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import gen_graph_golden as gg  # noqa: E402
import torch_tf  # noqa: E402
from e2e_asr_b200 import synth  # noqa: E402


def run_case(case):
    cname, enc_over, dec_over = gg.CASES[case]
    cfg = synth.get_config(cname)
    w = {k: v for k, v in synth.make_weights(cfg, bias_noise=0.1).items() if not k.startswith("model/ctc_")}
    batch = synth.make_batch(cfg)
    tf = torch_tf.make_tf(w)
    mods = gg.load_reference(tf)
    ep = mods["encoder"].Encoder.class_params()
    ep.hidden_size, ep.use_lstm, ep.out_prob = cfg.H, True, 1.0
    ep.update(enc_over)
    dp = mods["attn_decoder"].AttnDecoder.class_params()
    dp.hidden_size_dec, dp.emb_size, dp.vocab_size = cfg.Hd, cfg.E, cfg.V
    dp.attention_vec_size, dp.lm_hidden_size, dp.max_output = cfg.A, cfg.Hl, cfg.U
    dp.out_prob_dec, dp.samp_prob = 1.0, 0.0
    dp.update(dec_over)
    t = torch_tf.t
    with tf.variable_scope("model"):
        enc = mods["encoder"].Encoder(params=ep, isTraining=True)
        att, _, lens = enc(t(batch["logmel"].astype(np.float64)), t(batch["logmel_len"]), {"char": cfg.L})
        dec = mods["attn_decoder"].AttnDecoder(isTraining=True, params=dp, scope="char")
        dec_inp, seq_len = t(np.ascontiguousarray(batch["char"].T)), t(batch["char_len"])
        logits = dec(dec_inp, seq_len, att[cfg.L], lens[cfg.L])
        targets, _ = mods["tf_utils"].create_shifted_targets(dec_inp, seq_len)
        loss = mods["losses"].LossUtils.cross_entropy_loss(logits, targets, seq_len)
    loss.backward()
    out = {"loss": np.asarray(loss.detach())}
    for k, p in tf._graph.params.items():
        assert p.grad is not None, k
        out["grad/" + k] = p.grad.numpy()
    np.savez(os.path.join(HERE, "grad_%s.npz" % case), **out)
    return out


if __name__ == "__main__":
    for case in gg.CASES:
        o = run_case(case)
        print(case, "loss", float(o["loss"]), "grads", len(o) - 1,
              "max |g|", max(float(np.abs(v).max()) for k, v in o.items() if k != "loss"))

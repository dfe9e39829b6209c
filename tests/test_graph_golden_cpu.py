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

"""The oracle against outputs of the reference's own graph-building functions executed with a NumPy stand-in for
their TensorFlow ops (tests/golden/gen_tf_composition_golden.py): get_batch (frame stacking, time-major ids, eval
lengths), _get_pyramid_input (zero-padding for odd maxima, ceil lengths), cross_entropy_loss (mask, per-example
normalisation, batch mean) and create_shifted_targets."""
import os

import numpy as np

from oracle import model as om

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tf_composition.npz"))


def test_get_batch_frame_stacking_and_layout():
    logmel, char = G["get_batch/logmel"], G["get_batch/char"]
    for stack in (1, 3):
        for mode in ("train", "eval"):
            key = "get_batch/stack%d/%s/" % (stack, mode)
            assert np.array_equal(om.stack_frames(logmel.astype(np.float64), stack), G[key + "enc_in"])
            assert np.array_equal(char.T, G[key + "dec_in"])               # time-major ids (seq2seq_model.py:189)
            assert np.array_equal(G[key + "enc_len"], [7, 4, 1])
    assert np.array_equal(G["get_batch/stack1/train/dec_len"], [5, 2, 1])
    assert np.array_equal(G["get_batch/stack1/eval/dec_len"], [11, 11, 11])  # eval: len := max_output (:191-193)


def test_pyramid_padding_reshape_and_lengths():
    for name in ("odd", "even", "odd_short"):
        x, lens = G["pyramid/%s/x" % name].astype(np.float64), G["pyramid/%s/lens" % name]
        y, new_len, _ = om.pyramid_fwd(x, lens, 2)
        assert y.shape == G["pyramid/%s/y" % name].shape
        assert np.array_equal(y, G["pyramid/%s/y" % name])
        assert np.array_equal(new_len, G["pyramid/%s/new_len" % name])


def test_cross_entropy_loss_and_shifted_targets():
    logits, dec_inp, seq_len = G["loss/logits"].astype(np.float64), G["loss/dec_inp"], G["loss/seq_len"]
    U = int(seq_len.max())
    targets = dec_inp[1:U + 1]
    assert np.array_equal(targets, G["loss/targets"])
    mask = (np.arange(U)[:, None] < seq_len[None, :]).astype(np.float32).reshape(-1)
    assert np.array_equal(mask, G["loss/weights"])
    loss, dlogits = om.cross_entropy_loss(logits, targets, seq_len)
    assert abs(loss - float(G["loss/value"])) < 2e-6                      # the golden was reduced in float32
    # the product's host-side target shifting (tf_utils.create_shifted_targets restated on torch tensors)
    import torch
    from e2e_asr_b200.tf_utils import create_shifted_targets
    t2, w2 = create_shifted_targets(torch.from_numpy(dec_inp), torch.from_numpy(seq_len))
    assert np.array_equal(t2.numpy(), G["loss/targets"]) and np.array_equal(w2.numpy(), G["loss/weights"])

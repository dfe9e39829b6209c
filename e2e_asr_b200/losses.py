"""LossUtils (reference losses.py:6-35) plus the auxiliary CTC the north-star adds."""
import torch

from . import ops


class LossUtils(object):

    @staticmethod
    def cross_entropy_loss(logits, targets, seq_len_target):
        """Masked, per-example length-normalised sparse softmax CE, batch mean.

        logits: [(T*B), V] time-major rows; targets: [T, B] int64 (rows beyond the
        logits' T are ignored); seq_len_target: [B].  (losses.py:7-35)
        """
        B = targets.shape[1]
        U = logits.shape[0] // B
        targets = targets[:U]
        lens = ops.to_i32(seq_len_target, logits.device)
        return ops.CrossEntropyFn.apply(logits, targets, lens)

    @staticmethod
    def ctc_head_loss(states, kernel, bias, seq_len, labels, label_len, stash=None, max_label_len=None):
        """Builder-defined auxiliary CTC (no reference code exists):
        projection + tf.nn.ctc_loss semantics (blank = last class), batch mean.
        states: [B,T,D] or the time-major [T,B,D] view the encoder keeps for the
        "state" task (encoder.py:143-144,160-161)."""
        dev = states.device
        if max_label_len is None:          # sizes the extended-label state space; any bound >= the longest label works
            lab_host = ops.host_array(label_len)
            max_label_len = int(lab_host.max()) if len(lab_host) else 0
        max_l = int(max_label_len)
        return ops.CTCHeadFn.apply(states, kernel, bias, ops.to_i32(seq_len, dev), labels,
                                   ops.to_i32(label_len, dev), max_l, stash)

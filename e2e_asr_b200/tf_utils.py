"""create_shifted_targets (reference tf_utils.py:4-12). The checkpoint helpers of
tf_utils.py:17-90 are TF checkpoint I/O and out of scope."""
import torch


def create_shifted_targets(dec_input, seq_len):
    """Shift the dec_input by 1 to create the targets; also the [T*B] loss mask.
    dec_input: [U+1, B] ids (time major); seq_len: [B]."""
    targets = dec_input[1:]
    T = targets.shape[0]
    mask = (torch.arange(T, device=seq_len.device)[:, None] < seq_len[None, :]).to(torch.float32)
    return targets, mask.reshape(-1)

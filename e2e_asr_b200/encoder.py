"""Encoder: pyramidal multi-layer BiLSTM over log-mel frames.

Same constructor / __call__ signature, hyper-parameters and return values as the
reference `Encoder` (encoder.py:15-200); the TF graph ops are replaced by the
persistent-recurrence and GEMM kernels of libe2e_asr_b200.so.

Layout: every layer works on a zero-padded batch-major buffer [B, Tp_l, C] with
Tp_1 a multiple of 2^(#pyramid steps) and Tp_l >= max(len_l) + 1.  The pyramid
reshape (encoder.py:94-119) is then a free view, and the public states are
time-narrowed views with exactly the reference's shapes [B, T_l, 2H].
"""
import math

import numpy as np
import torch

from . import ops
from .base_params import BaseParams, Bunch
from .variables import default_store


class Encoder(BaseParams):
    """Encoder class that encodes input sequence."""

    @classmethod
    def class_params(cls):
        """Encoder class parameters (encoder.py:19-31)."""
        params = Bunch()
        params['bi_dir'] = True
        params['hidden_size'] = 256
        params['out_prob'] = 0.9
        params['skip_step'] = 2  # Pyramidal architecture
        params['initial_res_fac'] = 1
        params['use_lstm'] = False
        params['stack_cons'] = 1
        params['max_scaling_down'] = 8
        return params

    def __init__(self, params=None, isTraining=True, variables=None):
        self.params = params if params is not None else self.class_params()
        self.isTraining = isTraining
        self.variables = variables

    def _check_supported(self):
        p = self.params
        if not p.use_lstm and p.hidden_size > 512:
            raise NotImplementedError("Encoder: GRU cells (use_lstm=False) are built for hidden_size <= 512")
        if not (0.0 < p.out_prob <= 1.0):
            raise ValueError("Encoder: out_prob=%g must be in (0, 1]" % p.out_prob)
        if p.skip_step not in (1, 2):
            raise NotImplementedError("Encoder: skip_step must be 1 or 2")

    def _layer_vars(self, layer_depth, in_size):
        """Variables of RNNLayer<d> (encoder.py:72-81): U(-0.075, 0.075) kernels, zero biases."""
        vs = self.variables if self.variables is not None else default_store()
        H = self.params.hidden_size
        out = []
        if self.params.bi_dir:
            scopes = ["model/encoder/RNNLayer%d/bidirectional_rnn/%s/" % (layer_depth, d) for d in ("fw", "bw")]
        else:   # tf.nn.dynamic_rnn(..., scope=str(layer_depth)) inside the RNNLayer<d> scope (encoder.py:86-89)
            scopes = ["model/encoder/RNNLayer%d/%d/" % (layer_depth, layer_depth)]
        if not self.params.use_lstm:
            # tf.nn.rnn_cell.GRUCell: gates/{kernel,bias} (bias initialised to 1.0), candidate/{kernel,bias}
            for sc in scopes:
                out.append(vs.get(sc + "gru_cell/gates/kernel", (in_size + H, 2 * H), ("uniform", 0.075)))
                out.append(vs.get(sc + "gru_cell/gates/bias", (2 * H,), ("ones",)))
                out.append(vs.get(sc + "gru_cell/candidate/kernel", (in_size + H, H), ("uniform", 0.075)))
                out.append(vs.get(sc + "gru_cell/candidate/bias", (H,), ("zeros",)))
            return out
        for sc in scopes:
            out.append(vs.get(sc + "basic_lstm_cell/kernel", (in_size + H, 4 * H), ("uniform", 0.075)))
            out.append(vs.get(sc + "basic_lstm_cell/bias", (4 * H,), ("zeros",)))
        return out + [None] * (4 - len(out))

    def _layer_encoder_input(self, x_padded, lens_i32, max_len, layer_depth=1):
        """Run one (Bi)LSTM layer on a padded batch-major buffer (encoder.py:55-91)."""
        if self.params.use_lstm:
            k_fw, b_fw, k_bw, b_bw = self._layer_vars(layer_depth, x_padded.shape[2])
            out = ops.BiLSTMLayerFn.apply(x_padded, k_fw, b_fw, k_bw, b_bw, lens_i32, max_len)
        else:
            out = ops.BiGRULayerFn.apply(x_padded, lens_i32, max_len, *self._layer_vars(layer_depth, x_padded.shape[2]))
        if self.isTraining and self.params.out_prob < 1.0:
            # DropoutWrapper(cell, output_keep_prob=out_prob) iff training (encoder.py:49-52): the per-step outputs
            # of both directions are dropped, the recurrent state is not.  Mask stream = layer depth.
            out = ops.DropoutFn.apply(out, self.params.out_prob, self.dropout_seed, layer_depth)
        return out

    def __call__(self, encoder_input, seq_len, num_layers):
        """Run the encoder (encoder.py:122-180).

        Args:
            encoder_input: [B, T, F] float32 (batch major, zero padded past seq_len).
            seq_len: [B] valid frames per utterance.
            num_layers: dict task -> encoder depth; a task named "state" (or ending
                in "_ctc") gets the time-major view (encoder.py:143-144).
        Returns:
            attention_states {depth: [B,T_d,2H]}, time_major_states {depth: [T_d,B,2H]},
            seq_len_inps {depth: [B] int64}.
        """
        self._check_supported()
        params = self.params
        attention_states, time_major_states, seq_len_inps = {}, {}, {}
        max_depth = 0
        for task, num_layer in num_layers.items():
            if task == "state" or task.endswith("_ctc"):
                time_major_states[num_layer] = None
            else:
                attention_states[num_layer] = None
            max_depth = max(max_depth, num_layer)

        dev = encoder_input.device
        self.layer_done = {}
        if not hasattr(self, "dropout_seed"):
            self.dropout_seed = 0
        lens_host = np.asarray(ops.host_array(seq_len), np.int64)
        B, T, F = encoder_input.shape
        res = params.initial_res_fac
        # per-layer lengths are derived ON THE DEVICE from the batch's length tensor (a handful of tiny kernels), so a
        # captured step (GraphedStep) follows the lengths of the batch it is replayed on; the host copy only feeds
        # shape decisions (the maximum length)
        lens_dev = ops.to_i32(seq_len, dev)
        if res > 1:
            lens_host = -(-lens_host // res)
            lens_dev = (lens_dev + (res - 1)) // res
            T = -(-T // res)
        # number of pyramid reductions that will be applied (encoder.py:172)
        n_red, r = 0, res
        for i in range(max_depth):
            if params.skip_step > 1 and i != max_depth - 1 and r < params.max_scaling_down:
                n_red += 1
                r *= params.skip_step
        q = 2 ** n_red
        Tp = q * (-(-T // q) + 1)
        x = ops.prepare_input(encoder_input, Tp, params.stack_cons, max(res, 1))
        if self.isTraining is False:
            x = x.detach()

        if params.use_lstm and x.is_cuda and max_depth > 1:
            # the weights of the layers above the first are packed on a side stream while layer 1 runs
            pre, in_size, r = [], x.shape[2], res
            nd = 2 if params.bi_dir else 1
            for i in range(max_depth):
                if i > 0:
                    kv = self._layer_vars(i + 1, in_size)
                    pre.append(([kv[0], kv[2]][:nd], [kv[1], kv[3]][:nd], in_size, params.hidden_size))
                in_size = nd * params.hidden_size
                if params.skip_step > 1 and i != (max_depth - 1) and r < params.max_scaling_down:
                    in_size *= 2
                    r *= params.skip_step
            ops.prepack_lstm(pre, dev)

        T_l = T
        for i in range(max_depth):
            layer_depth = i + 1
            # loop bound of the recurrence: the longest utterance of the batch, or -- for a step captured once per
            # length bucket (GraphedStep(bucket=True)) -- the padded shape, the extra steps being masked
            max_len = (T_l if getattr(self, "shape_bounds", False) else int(lens_host.max())) if B else 0
            out = self._layer_encoder_input(x, lens_dev, max_len, layer_depth)       # [B, Tp, 2H]
            if out.is_cuda:
                # consumers on other streams (auxiliary heads) may start as soon as THIS layer is done
                ev = torch.cuda.Event()
                ev.record()
                self.layer_done[layer_depth] = ev
            view = out[:, :T_l]
            if layer_depth in time_major_states:
                time_major_states[layer_depth] = view.transpose(0, 1)
            if layer_depth in attention_states:
                attention_states[layer_depth] = view
            sl = lens_dev.to(torch.int64)
            sl._host = lens_host.copy()
            sl._i32 = lens_dev
            seq_len_inps[layer_depth] = sl
            if params.skip_step > 1 and i != (max_depth - 1) and res < params.max_scaling_down:
                # _get_pyramid_input: frame 2k || frame 2k+1, len = ceil(len/2)
                x = out.view(B, Tp // 2, out.shape[2] * 2)
                Tp //= 2
                T_l = -(-T_l // 2)
                lens_host = -(-lens_host // 2)
                lens_dev = (lens_dev + 1) // 2
                res *= params.skip_step
            else:
                x = out
        return attention_states, time_major_states, seq_len_inps

    @classmethod
    def add_parse_options(cls, parser):
        # flag names and defaults of encoder.py:183-200
        parser.add_argument("-out_prob", "--out_prob", default=0.9, type=float,
                            help="Output keep probability for dropout")
        parser.add_argument("-use_lstm", "--use_lstm", default=True, action="store_true", help="LSTM cells")
        parser.add_argument("-hsize", "--hidden_size", default=256, type=int, help="Hidden layer size")
        parser.add_argument("-skip_step", "--skip_step", default=2, type=int,
                            help="Frame skipping factor as we go up the stacked layers")
        parser.add_argument("-init_res_fac", "--initial_res_fac", default=1, type=int,
                            help="Initial resolution factor")
        parser.add_argument("-stack_cons", default=1, type=int, help="Stacking consecutive frames in input")
        parser.add_argument("-max_scaling_down", default=8, type=int, help="Maximum reduction in resolution")

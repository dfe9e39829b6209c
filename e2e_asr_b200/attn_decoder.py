"""Attention decoder with the reference's signature (attn_decoder.py:18-186):
`AttnDecoder(isTraining, params, scope)(decoder_inp, seq_len,
encoder_hidden_states, seq_len_inp)` -> logits [(U*B), V], time-major rows."""
import numpy as np
import torch

from . import ops
from .decoder import Decoder


class AttnDecoder(Decoder):
    """Implements the attention decoder of encoder-decoder framework."""

    @classmethod
    def class_params(cls):
        params = super(AttnDecoder, cls).class_params()
        params['attention_vec_size'] = 128
        params['lm_hidden_size'] = 256
        params['ind_softmax'] = False
        return params

    def __init__(self, isTraining, params=None, scope=None, variables=None):
        super(AttnDecoder, self).__init__(isTraining=isTraining, params=params, variables=variables)
        self.scope = scope
        self.stash = {}

    def scope_name(self):
        return "model/rnn_decoder" + ("" if self.scope is None else "_" + self.scope)

    def get_variables(self, attn_size):
        """All variables of the decoder scope, created with the reference's names,
        shapes and initialisers."""
        p, vs, s = self.params, self._store(), self.scope_name() + "/"
        E, Hd, Hl, A, V = p.emb_size, p.hidden_size_dec, p.lm_hidden_size, p.attention_vec_size, p.vocab_size
        D = attn_size
        out_name = "OutputProjection2" if p.ind_softmax else "OutputProjection"   # attn_decoder.py:119-125
        if self.general_cells():
            return self._general_variables(vs, s, E, Hd, Hl, A, V, D, out_name)
        v = dict(
            emb=vs.get(s + "decoder/embedding", (V, E), ("uniform", 1.0)),          # decoder.py:97-99
            attn_w=vs.get(s + "AttnW", (1, 1, D, A)),
            attn_v=vs.get(s + "AttnV", (A,)),
            lm_k=vs.get(s + "rnn/basic_lstm_cell/kernel", (E + Hl, 4 * Hl)),
            lm_b=vs.get(s + "rnn/basic_lstm_cell/bias", (4 * Hl,), ("zeros",)),
            dec_k=vs.get(s + "rnn/basic_lstm_cell_1/kernel", (E + Hd, 4 * Hd)),
            dec_b=vs.get(s + "rnn/basic_lstm_cell_1/bias", (4 * Hd,), ("zeros",)),
            q_k=vs.get(s + "rnn/Attention/kernel", (Hd, A)),
            q_b=vs.get(s + "rnn/Attention/bias", (A,), ("zeros",)),
            ap_k=vs.get(s + "rnn/AttnProjection/kernel", (Hd + D, Hd)),
            ap_b=vs.get(s + "rnn/AttnProjection/bias", (Hd,), ("zeros",)),
            out_k=vs.get(s + "rnn/%s/kernel" % out_name, (Hd, V)),
            out_b=vs.get(s + "rnn/%s/bias" % out_name, (V,), ("zeros",)),
            in_k=vs.get(s + "rnn/InputProjection/kernel", (Hd + D, E)),
            in_b=vs.get(s + "rnn/InputProjection/bias", (E,), ("zeros",)),
            sp_k=None, sp_b=None)
        if Hl != Hd:                                                                 # attn_decoder.py:149-151
            v["sp_k"] = vs.get(s + "rnn/SimpleProjection/kernel", (Hl, Hd))
            v["sp_b"] = vs.get(s + "rnn/SimpleProjection/bias", (Hd,), ("zeros",))
        return v

    def _general_variables(self, vs, s, E, Hd, Hl, A, V, D, out_name):
        """Variables when lm_cell / the decoder cell are MultiRNNCell stacks and / or GRU cells (decoder.py:49-72).
        lm_cell is called first inside the raw_rnn loop, the decoder cell second, so TF's scope uniquifier gives
        rnn/<cell>/... and rnn/<cell>_1/... (the single-cell names beam_search.py:56-98 reads), resp.
        rnn/multi_rnn_cell/cell_<l>/<cell>/... and rnn/multi_rnn_cell_1/cell_<l>/<cell>/... for stacks."""
        p = self.params
        nl, lstm = p.num_layers_dec, p.use_lstm
        cell = "basic_lstm_cell" if lstm else "gru_cell"
        stacks = []
        for sfx, Hc in (("", Hl), ("_1", Hd)):
            layers = []
            for l in range(nl):
                base = s + "rnn/" + (cell + sfx + "/" if nl == 1 else "multi_rnn_cell%s/cell_%d/%s/" % (sfx, l, cell))
                I_ = E if l == 0 else Hc
                if lstm:
                    layers.append((vs.get(base + "kernel", (I_ + Hc, 4 * Hc)), vs.get(base + "bias", (4 * Hc,), ("zeros",))))
                else:
                    layers.append((vs.get(base + "gates/kernel", (I_ + Hc, 2 * Hc)),
                                   vs.get(base + "gates/bias", (2 * Hc,), ("ones",)),
                                   vs.get(base + "candidate/kernel", (I_ + Hc, Hc)),
                                   vs.get(base + "candidate/bias", (Hc,), ("zeros",))))
            stacks.append(layers)
        v = dict(
            emb=vs.get(s + "decoder/embedding", (V, E), ("uniform", 1.0)),
            attn_w=vs.get(s + "AttnW", (1, 1, D, A)),
            attn_v=vs.get(s + "AttnV", (A,)),
            lm_cells=stacks[0], dec_cells=stacks[1],
            q_k=vs.get(s + "rnn/Attention/kernel", (Hd, A)),
            q_b=vs.get(s + "rnn/Attention/bias", (A,), ("zeros",)),
            ap_k=vs.get(s + "rnn/AttnProjection/kernel", (Hd + D, Hd)),
            ap_b=vs.get(s + "rnn/AttnProjection/bias", (Hd,), ("zeros",)),
            out_k=vs.get(s + "rnn/%s/kernel" % out_name, (Hd, V)),
            out_b=vs.get(s + "rnn/%s/bias" % out_name, (V,), ("zeros",)),
            in_k=vs.get(s + "rnn/InputProjection/kernel", (Hd + D, E)),
            in_b=vs.get(s + "rnn/InputProjection/bias", (E,), ("zeros",)),
            sp_k=None, sp_b=None)
        if Hl != Hd:
            v["sp_k"] = vs.get(s + "rnn/SimpleProjection/kernel", (Hl, Hd))
            v["sp_b"] = vs.get(s + "rnn/SimpleProjection/bias", (Hd,), ("zeros",))
        return v

    def __call__(self, decoder_inp, seq_len, encoder_hidden_states, seq_len_inp):
        """decoder_inp: [U+1, B] int64 ids (row 0 = GO); seq_len: [B] number of
        targets; encoder_hidden_states: [B, T_enc, D]; seq_len_inp: [B]."""
        self._check_supported()
        enc = encoder_hidden_states
        dev = enc.device
        v = self.get_variables(enc.shape[2])
        lens_host = np.asarray(ops.host_array(seq_len))
        U = int(lens_host.max()) if len(lens_host) else 0        # raw_rnn stops when all rows are finished
        if getattr(self, "shape_bounds", False):                 # bucket-captured step: all padded steps, masked
            U = int(decoder_inp.shape[0]) - 1
        lens = ops.to_i32(seq_len, dev)
        enc_len = ops.to_i32(seq_len_inp, dev)
        rule = self.input_rule()
        if self.general_cells():
            drop = None
            if self.isTraining and self.params.out_prob_dec < 1.0:
                drop = (self.params.out_prob_dec, getattr(self, "dropout_seed", 0), getattr(self, "dropout_stream", 0))
            cells = (enc, v, v["lm_cells"], v["dec_cells"], self.params.use_lstm)
            if rule == "greedy":        # eval mode: U = max_output steps for every row (seq2seq_model.py:191-193)
                with torch.no_grad():
                    return ops.attn_decoder_stepwise(*cells, decoder_inp, lens, enc_len, U, None, feedback="greedy")
            if rule == "sample":
                # scheduled sampling as for the single cell: realise the input ids without a tape (same Philox
                # streams), then take the teacher-forced step on them
                from .host_utils import philox_uniform
                seed, stream = getattr(self, "dropout_seed", 0), getattr(self, "dropout_stream", 0)
                use = [False] + [not (philox_uniform(t, 200 + stream, seed) < 1.0 - self.params.samp_prob)
                                 for t in range(1, U)]
                ids = decoder_inp[:U].clone()
                if any(use):
                    last = max(t for t in range(U) if use[t])
                    with torch.no_grad():
                        ops.attn_decoder_stepwise(*cells, decoder_inp, lens, enc_len, last, None, drop=drop,
                                                  feedback=dict(use_sample=use, seed=seed, offset=300 + stream, ids=ids))
                decoder_inp = self.stash["realized_ids"] = ids
            return ops.attn_decoder_stepwise(*cells, decoder_inp, lens, enc_len, U, self.stash, drop=drop)
        if rule in ("teacher", "sample"):
            # DropoutWrapper(output_keep_prob=out_prob_dec) iff training (decoder.py:60-63) acts on lm_cell's
            # output only: raw_loop_function reads the decoder cell through get_state(state) = state.c and never
            # uses cell_output (attn_decoder.py:114-118), so the decoder-LSTM wrapper has no effect on the result.
            lm_drop = None
            if self.isTraining and self.params.out_prob_dec < 1.0:
                lm_drop = (self.params.out_prob_dec, getattr(self, "dropout_seed", 0),
                           100 + getattr(self, "dropout_stream", 0))
            samp = None
            if rule == "sample":
                from .host_utils import philox_uniform
                seed, stream = getattr(self, "dropout_seed", 0), getattr(self, "dropout_stream", 0)
                if (ops._DECODER_IMPL == "persist" and ops._persist_fits(enc, v["dec_k"], v["q_k"], U)
                        and v["lm_k"].shape[1] // 4 in (128, 256)):       # the carried-state LM-LSTM entry
                    # scheduled sampling inside the training forward pass: the persistent loop runs in segments between
                    # the sampled steps (ops.AttnDecoderFnV2._forward_sampled); one scalar draw per step decides
                    use = [False] + [not (philox_uniform(t, 200 + stream, seed) < 1.0 - self.params.samp_prob)
                                     for t in range(1, U)]
                    samp = (use, seed, 300 + stream)
                else:
                    # shapes the persistent kernels cannot hold: realise the input ids without a tape (one decoder step
                    # at a time), then take the teacher-forced step on them
                    from .inference import sample_decode_ids
                    with torch.no_grad():
                        ids = sample_decode_ids(v, decoder_inp, lens, U, enc.detach(), enc_len, self.params.samp_prob,
                                                seed, stream, lm_drop=lm_drop)
                    decoder_inp = self.stash["realized_ids"] = ids
            # teacher-forced ids are known when the step starts: the LM side may run ahead of the encoder
            self.stash["early_lm"] = rule == "teacher"
            return ops.attn_decoder_apply(
                enc, v["emb"], v["attn_w"], v["attn_v"], v["lm_k"], v["lm_b"], v["dec_k"], v["dec_b"], v["q_k"],
                v["q_b"], v["ap_k"], v["ap_b"], v["out_k"], v["out_b"], v["in_k"], v["in_b"], v["sp_k"], v["sp_b"],
                decoder_inp, lens, enc_len, U, self.stash, lm_drop=lm_drop, samp=samp)
        from .inference import greedy_decode_logits
        with torch.no_grad():
            return greedy_decode_logits(v, decoder_inp, lens, U, enc, enc_len)

    @classmethod
    def add_parse_options(cls, parser):
        # flag names and defaults of attn_decoder.py:174-186
        super(AttnDecoder, cls).add_parse_options(parser)
        parser.add_argument("-samp_prob", "--samp_prob", default=0.1, type=float,
                            help="Scheduled sampling probability")
        parser.add_argument("-attn_vec_size", "--attention_vec_size", default=128, type=int,
                            help="Attention vector size")
        parser.add_argument("-lm_hsize", "--lm_hidden_size", default=256, type=int, help="Hidden Size of LM layer")
        parser.add_argument('-ind_softmax', "--ind_softmax", default=False, action="store_true",
                            help="Independent (from LM) softmax params")

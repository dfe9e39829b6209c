"""Decoder base class (reference decoder.py:17-193): hyper-parameters, variable
creation and the choice of next-input rule.  The actual stepping lives in
attn_decoder.py / the CUDA decoder loop."""
from .base_params import BaseParams, Bunch
from .variables import default_store


class Decoder(BaseParams):
    """Base class for decoder in encoder-decoder framework."""

    @classmethod
    def class_params(cls):
        """Decoder class parameters (decoder.py:22-34)."""
        params = Bunch()
        params['out_prob_dec'] = 0.9
        params['hidden_size_dec'] = 256
        params['num_layers_dec'] = 1
        params['emb_size'] = 256
        params['vocab_size'] = 1000
        params['samp_prob'] = 0.1
        params['max_output'] = 400
        params['use_lstm'] = True
        return params

    def __init__(self, isTraining=True, params=None, variables=None):
        self.params = self.class_params() if params is None else params
        self.isTraining = isTraining
        self.variables = variables

    def _store(self):
        return self.variables if self.variables is not None else default_store()

    def _check_supported(self):
        p = self.params
        if p.num_layers_dec < 1:
            raise ValueError("Decoder: num_layers_dec=%d must be >= 1" % p.num_layers_dec)
        if self.general_cells():
            # MultiRNNCell stacks / GRU cells run the step-by-step path (ops.attn_decoder_stepwise) for every input
            # rule: teacher forcing, scheduled sampling, eval-mode greedy feedback; output dropout
            if self.isTraining and p.out_prob_dec < 1.0 and (p.lm_hidden_size % 4 or p.hidden_size_dec % 4):
                raise NotImplementedError("Decoder: dropout needs hidden sizes that are multiples of 4")
        if not (0.0 < p.out_prob_dec <= 1.0):
            raise ValueError("Decoder: out_prob_dec=%g must be in (0, 1]" % p.out_prob_dec)
        if not (0.0 <= p.samp_prob <= 1.0):
            raise ValueError("Decoder: samp_prob=%g must be in [0, 1]" % p.samp_prob)

    def __call__(self, decoder_inp, seq_len, encoder_hidden_states, seq_len_inp):
        """Abstract in the reference too (decoder.py:117-137): implemented by AttnDecoder."""
        raise NotImplementedError("Decoder.__call__ is abstract: use AttnDecoder")

    def general_cells(self):
        """True when lm_cell / the decoder cell are not single LSTM cells (decoder.py:49-72)."""
        return self.params.num_layers_dec > 1 or not self.params.use_lstm

    def get_state(self, state):
        """The attention query / projection input is the LSTM CELL state c of the
        last layer (decoder.py:74-82). `state` is a (c, h) pair."""
        if self.params.num_layers_dec > 1:
            state = state[-1]
        return state[0] if self.params.use_lstm else state

    def input_rule(self):
        """Which token feeds step t+1 (decoder.py:103-115): 'teacher' in training
        with samp_prob == 0, 'sample' with scheduled sampling, 'greedy' at eval."""
        if self.isTraining:
            return "sample" if self.params.samp_prob > 0 else "teacher"
        return "greedy"

    @classmethod
    def add_parse_options(cls, parser):
        # flag names and defaults of decoder.py:182-193
        parser.add_argument("-hsize_dec", "--hidden_size_dec", default=256, type=int,
                            help="Hidden size of decoder RNN")
        parser.add_argument("-emb_size", "--emb_size", default=256, type=int, help="Embedding size")
        parser.add_argument("-num_layers_dec", "--num_layers_dec", default=1, type=int,
                            help="Number of RNN layers")
        parser.add_argument("-out_prob_dec", "--out_prob_dec", default=0.9, type=float, help="1 - dropout_prob")

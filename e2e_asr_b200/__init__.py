"""e2e_asr_b200: B200-native (sm_100a) implementation of the training hot path of
shtoshni/e2e_asr -- pyramidal BiLSTM encoder, attention decoder, sequence CE and
auxiliary CTC -- behind the reference's Python class API.

Importing the package is CPU-safe; using any op requires the in-tree
libe2e_asr_b200.so (build with `__graft_entry__.build()`) and a CUDA device.
"""
from .base_params import BaseParams, Bunch  # noqa: F401
from .data_utils import EOS_ID, GO_ID, PAD_ID  # noqa: F401
from .host_utils import BasicLSTM, BeamEntry  # noqa: F401


def __getattr__(name):
    # torch-dependent classes are imported lazily so `import e2e_asr_b200` stays light
    import importlib
    table = {"Encoder": ".encoder", "Decoder": ".decoder", "AttnDecoder": ".attn_decoder",
             "LossUtils": ".losses", "Seq2SeqModel": ".seq2seq_model", "BeamSearch": ".beam_search",
             "VariableStore": ".variables", "SpeechDataset": ".speech_dataset"}
    if name in table:
        return getattr(importlib.import_module(table[name], __name__), name)
    raise AttributeError(name)

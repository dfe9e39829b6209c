"""One eager 256-utterance beam decode (ncu target)."""
import sys; sys.path.insert(0, '.')
import torch
from e2e_asr_b200 import synth
from e2e_asr_b200.beam_search import BeamSearch
cfg = synth.get_config("cfg2")
w = synth.make_weights(cfg)
encs = synth.make_beam_eval_batch(cfg, 256)
sp = BeamSearch.class_params(); sp.beam_size = 10
bs = BeamSearch(w, sp, device="cuda:0")
bs.MAX_STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bs.decode_batch(encs, use_graph=False)
torch.cuda.synchronize()

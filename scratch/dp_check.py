"""2-GPU check of the data-parallel step (run under torchrun): the clipped gradient of the DP step (per-rank shards,
overlapped span all-reduce, 1/n folded into the clip) must equal the single-GPU step on the GLOBAL batch."""
import os, sys, faulthandler
sys.path.insert(0, '.')
faulthandler.enable()
faulthandler.dump_traceback_later(int(os.environ.get("DP_CHECK_DUMP_S", "70")), exit=True)    # a hang prints every thread's stack
import numpy as np, torch, torch.distributed as dist
from e2e_asr_b200 import ops, synth
from e2e_asr_b200 import dist as edist
from e2e_asr_b200.testing import build_model

def say(*a):
    print("[rank %s]" % os.environ.get("RANK", "0"), *a, flush=True)


rank, world, local = edist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
ops.set_gemm_mode("tf32x3")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
cfg_g = synth.get_config(name, B=8) if name == "cfg1" else synth.get_config(name)
w = synth.make_weights(cfg_g)
gbatch = synth.make_batch(cfg_g)
shard = edist.shard_batch(gbatch, rank, world)
cfg_r = synth.get_config(name, B=cfg_g.B // world)
warm = torch.zeros(1, device=dev)
dist.all_reduce(warm)
torch.cuda.synchronize()
say("communicator up")
red = edist.GradAllReducer()
model = build_model(cfg_r, w, device=dev, reducer=red)
results = {}
for mode in ("eager", "graph"):
    say("mode", mode)
    if mode == "graph":
        gs = model.graphed_step(shard)
        say("captured")
        gs.step(shard); gs.step(shard)
    else:
        model.run_step(shard)
        torch.cuda.synchronize()
        say("first eager step done")
        model.run_step(shard)
    torch.cuda.synchronize()
    results[mode] = (model.variables.flat_grads().detach().clone(), float(model.grad_norm))
    print("rank %d %s: norm %.6f collectives/step %d" % (rank, mode, results[mode][1], red.collectives), flush=True)
    red.collectives = 0
if rank == 0:
    ref = build_model(cfg_g, w, device=dev)
    ref.run_step(gbatch)
    torch.cuda.synchronize()
    g = ref.variables.flat_grads()
    gm = float(g.abs().max())
    for mode, (gd, nrm) in results.items():
        print("%s: global norm %.6f vs dp %.6f; max |dg| / max|g| = %.3e" %
              (mode, float(ref.grad_norm), nrm, float((gd - g).abs().max()) / gm), flush=True)
dist.barrier()
# timing: K replays
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(K):
    gs.step()
e1.record(); torch.cuda.synchronize()
print("rank %d: %.3f ms/step (graph, dp%d)" % (rank, e0.elapsed_time(e1) / K, world), flush=True)
# a CUDA graph holding NCCL nodes must be gone before the process group is destroyed (destroy blocks forever otherwise)
dist.barrier()
torch.cuda.synchronize()
del gs
del model
import gc
gc.collect()
torch.cuda.synchronize()
say("graph released")
dist.destroy_process_group()
say("process group destroyed")

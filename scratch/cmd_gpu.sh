for rep in 1 2 3; do for kb in 0 96 256; do
E2E_MAX_KBLOCKS=$kb timeout 300 python bench.py --steps 20 --warmup 5 --no-beam --no-cpu-baseline > gpurun_out/r2x_$kb.json 2> gpurun_out/r2x.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2x_$kb.json").read().strip().splitlines()[-1]); b=d["breakdown"]
print("kb=$kb", round(d["ms_per_step"],3), round(d["e2e"]["value"]), {k:round(b[k]["main_stream_ms_per_step"],3) for k in ("enc_rec_bwd","enc_rec_fwd")}, round(b["e2e_gemm"]["ms_per_step"],3))
P
done; done

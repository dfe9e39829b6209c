for rep in 1 2; do for gs in 0 1; do
E2E_SPLIT_GRID_STRIDE=$gs timeout 300 python bench.py --steps 20 --warmup 5 --no-beam --no-cpu-baseline > gpurun_out/r3a_$gs.json 2> gpurun_out/r3a.err
python - <<P
import json
d=json.loads(open("gpurun_out/r3a_$gs.json").read().strip().splitlines()[-1]); b=d["breakdown"]
print("gridstride=$gs", round(d["ms_per_step"],3), round(d["e2e"]["value"]), {k:round(b[k]["main_stream_ms_per_step"],3) for k in ("enc_rec_bwd","enc_rec_fwd","e2e_gemm")}, round(b["e2e_split_lo"]["ms_per_step"],3))
P
done; done

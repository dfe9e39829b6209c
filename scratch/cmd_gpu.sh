timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py -q -x -k "not beam" 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-beam --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2y_bench.json").read().strip().splitlines()[-1]); b=d["breakdown"]
print(round(d["ms_per_step"],3), round(d["e2e"]["value"]), {k:round(b[k]["main_stream_ms_per_step"],3) for k in ("enc_rec_bwd","enc_rec_fwd","e2e_lstm_pack_weights")})
P
done

timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/r2s_full.log 2>&1; echo "suite rc=$?"; grep -v "^frame" gpurun_out/r2s_full.log | tail -4 | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("pct_of_roofline"))
b=d.get("beam_decode",{}); print({k:b.get(k) for k in ("value","first_call_utt_s","capture_call_utt_s","ids_equal","ms_per_decoding_step")})
print(d.get("cpu_baseline"))
P

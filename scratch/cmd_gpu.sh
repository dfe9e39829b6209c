timeout 600 python -m pytest tests/test_gpu_beam.py tests/test_gpu_fullsize.py -q -x -k "beam" 2>&1 | tail -5
PYTHONPATH=. timeout 300 python scratch/time_beam.py 2>&1 | tail -14
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python - <<P
import json
d=json.loads(open("gpurun_out/r2r_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"])
print(json.dumps(d.get("beam_decode"))[:900])
P

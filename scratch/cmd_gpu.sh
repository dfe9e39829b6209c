timeout 1000 python -m pytest tests -m gpu -q > gpurun_out/r3e_full.log 2>&1; echo "suite rc=$?"; grep -v "^frame" gpurun_out/r3e_full.log | tail -2 | cut -c1-200
timeout 400 python bench.py --steps 10 --warmup 3 --no-beam --no-cpu-baseline > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r3e_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

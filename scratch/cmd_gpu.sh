run() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r3c_$name.json 2> gpurun_out/r3c_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/r3c_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  ', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']))"; }
run default --steps 20 --warmup 5
run tf32x3 --steps 10 --warmup 3 --gemm tf32x3 --no-beam --no-cpu-baseline
run cfg4 --steps 10 --warmup 3 --config cfg4 --no-beam --no-cpu-baseline
run cfg5 --steps 5 --warmup 3 --config cfg5 --no-beam --no-cpu-baseline
run dropout --steps 10 --warmup 3 --dropout --no-beam --no-cpu-baseline
run defaults --steps 10 --warmup 3 --defaults --no-beam --no-cpu-baseline

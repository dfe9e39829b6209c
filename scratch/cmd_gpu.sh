timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r3f_n4.json 2> gpurun_out/r3f_n4.err; echo "rc=$?"
tail -1 gpurun_out/r3f_n4.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"

#!/bin/bash
# usage (under gpurun): bash scratch/profile_round.sh <tag>
# plain run first (must exit 0), then the ncu launch list and the full captures of the top kernels; the ncu passes
# time the step kernel by kernel (--no-graph): a graph replay runs the identical kernels
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-beam --no-graph"
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-gemm_tc_kernel rec_fwd_ws_kernel dec_fwd_persist_kernel}; do
  SKIP=0; [ "$K" = gemm_tc_kernel ] && SKIP=40
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*$TAG*

#!/bin/bash
# usage (under gpurun): bash profiles/profile_round.sh <tag>
# plain run first (must exit 0), then the ncu launch list and the full captures of the top kernels; the ncu passes
# time the step kernel by kernel (--no-graph): a graph replay runs the identical kernels.  Read the reports back here with
#   python profiles/summarize_launches.py gpurun_out/launches_<tag>.csv > profiles/<tag>_launches_summary.txt
#   python profiles/extract_ncu.py gpurun_out/prof_*_<tag>.ncu-rep > profiles/<tag>_ncu_full_summary.txt
#   python profiles/extract_ncu.py --json profiles/ncu_traffic.json gpurun_out/prof_*_<tag>.ncu-rep
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-beam --no-graph"
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-rec_fwd_ws_kernel rec_bwd_ws_kernel gemm_tc_kernel dec_fwd_persist_kernel dec_bwd_persist_kernel ctc_sweep_kernel}; do
  SKIP=0; [ "$K" = gemm_tc_kernel ] && SKIP=40
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 2 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  echo "$K rc=$?"
done
ls -la gpurun_out/*$TAG*

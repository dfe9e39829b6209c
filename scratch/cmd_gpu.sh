timeout 200 python -m pytest tests/test_gpu_kernels.py -q -k "rows_of_very" > gpurun_out/r2o_rows.log 2>&1; grep -v "^frame" gpurun_out/r2o_rows.log | grep -E "Error|assert|passed|failed" | head -20
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -x -k "bilstm or fullsize_step" 2>&1 | tail -3
for i in 1 2 3 4 5 6; do timeout 200 python -m pytest tests/test_gpu_fullsize.py -q -x -k "cfg5 and f16x2" > gpurun_out/r2o_t$i.log 2>&1; echo "run $i rc=$?"; done

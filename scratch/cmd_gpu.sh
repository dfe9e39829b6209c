timeout 600 python -m pytest tests/test_gpu_beam.py tests/test_gpu_fullsize.py -q -x -k "beam or gemm_f64" 2>&1 | tail -4
PYTHONPATH=. timeout 300 python scratch/time_beam.py 2>&1 | tail -12

sed -i 's/for B in \[16, 32, 48, 56, 64, 112, 128, 256\]:/for B in [112]:/' scratch/time_rec512.py
PYTHONPATH=. timeout 300 python scratch/time_rec512.py 2>&1 | tail -2
PYTHONPATH=. timeout 400 ncu --set full --clock-control none --import-source on -k regex:rec_fwd_h512 -c 1 -f -o gpurun_out/prof_rec_fwd_h512_r2f python scratch/time_rec512.py > gpurun_out/ncu_h512f.log 2>&1; echo "rc=$?"
PYTHONPATH=. timeout 400 ncu --set full --clock-control none --import-source on -k regex:rec_bwd_h512 -c 1 -f -o gpurun_out/prof_rec_bwd_h512_r2f python scratch/time_rec512.py > gpurun_out/ncu_h512b.log 2>&1; echo "rc=$?"

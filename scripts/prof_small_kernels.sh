set -x
python scripts/time_small.py > gpurun_out/r02b_time_small.txt 2>&1
head -8 gpurun_out/r02b_time_small.txt
ncu --set full --import-source on --clock-control none -k regex:"chol_inv_kernel" -c 1 -f -o gpurun_out/r02b_chol python scripts/time_small.py > gpurun_out/r02b_ncu_small.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"syev_chol_jacobi_kernel" -c 2 -f -o gpurun_out/r02b_syev python scripts/time_small.py >> gpurun_out/r02b_ncu_small.log 2>&1
tail -3 gpurun_out/r02b_ncu_small.log

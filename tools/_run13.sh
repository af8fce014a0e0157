mkdir -p gpurun_out
python -m pytest tests -x -q -s -m gpu > gpurun_out/r02_pytest_gpu_c.txt 2>&1
tail -3 gpurun_out/r02_pytest_gpu_c.txt
grep "saturated COMAN\|state-input" gpurun_out/r02_pytest_gpu_c.txt | cut -c1-500
python tools/shape_sweep.py --steps 30 --seed 2026 --prec f64 2>&1 | cut -c1-230
python tools/states_profile.py | cut -c1-100
python tools/diff_kernels.py > gpurun_out/r02_diff_kernels.txt 2>&1; tail -2 gpurun_out/r02_diff_kernels.txt; grep -vc "status mismatches 0" gpurun_out/r02_diff_kernels.txt

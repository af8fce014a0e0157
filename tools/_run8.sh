set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -s -m gpu > gpurun_out/r02_pytest_gpu_b.txt 2>&1
tail -4 gpurun_out/r02_pytest_gpu_b.txt
grep "Cartesian/CoM\|state-input\|logged level\|fp32 C\|vs live" gpurun_out/r02_pytest_gpu_b.txt | cut -c1-400
python tools/diff_kernels.py > gpurun_out/r02_diff_kernels.txt 2>&1; tail -4 gpurun_out/r02_diff_kernels.txt

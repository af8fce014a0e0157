export LD_LIBRARY_PATH=$PWD/opensot_b200/lib:$PWD/oracle/_ref
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02d_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r02d_pytest_gpu.txt
python tests/gpu_prof.py 4096 > gpurun_out/r02d_prof_4096.txt 2>&1
python tests/gpu_prof.py 64 > gpurun_out/r02d_prof_64.txt 2>&1
sh tools/evidence_r02_ncu.sh a
ls -la gpurun_out/*.ncu-rep

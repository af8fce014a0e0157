mkdir -p gpurun_out
python tools/aux_kernels_bench.py | tee gpurun_out/r02_aux_kernels.jsonl
python tools/states_profile.py | cut -c1-120
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kinematics.py -x -q -m gpu 2>&1 | tail -2

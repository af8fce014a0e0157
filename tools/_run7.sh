set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kinematics.py tests/test_gpu_host_cpp.py -x -q -s -m gpu > gpurun_out/r02_pytest_kin.txt 2>&1
tail -4 gpurun_out/r02_pytest_kin.txt; grep "Cartesian/CoM\|state-input" gpurun_out/r02_pytest_kin.txt
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_states.json 2>gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r02_bench_states.json').read().strip().splitlines()[-1])
print(l['ms_per_step'], l['e2e'], l.get('e2e_states'))"

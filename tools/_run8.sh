set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -s -m gpu > gpurun_out/r02_pytest_gpu_b.txt 2>&1
tail -4 gpurun_out/r02_pytest_gpu_b.txt
grep "Cartesian/CoM\|state-input\|logged level\|fp32 C\|vs live" gpurun_out/r02_pytest_gpu_b.txt | cut -c1-400
python bench.py --no-cpu-baseline > gpurun_out/r02_bench_states.json 2>gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r02_bench_states.json').read().strip().splitlines()[-1])
print(l['ms_per_step'], l['unsolved_instances'], l['e2e'], l.get('e2e_states'))"
python tools/diff_kernels.py 2>&1 | tail -5

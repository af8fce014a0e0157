set -x
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/opensot_b200/lib:$PWD/oracle/_ref
python -m pytest tests/test_gpu_host_cpp.py -x -q -s > gpurun_out/r02_pytest_host_cpp.txt 2>&1
tail -3 gpurun_out/r02_pytest_host_cpp.txt
./tests/cpp/_bin/bench_dropin C29 4096 0.1 31 10 2>/dev/null | grep "^{" > gpurun_out/r02_bench_dropin.jsonl
./tests/cpp/_bin/bench_dropin C38c 4096 0.1 31 5 2>/dev/null | grep "^{" >> gpurun_out/r02_bench_dropin.jsonl
./tests/cpp/_bin/bench_dropin C35f 4096 0.1 31 5 2>/dev/null | grep "^{" >> gpurun_out/r02_bench_dropin.jsonl
cat gpurun_out/r02_bench_dropin.jsonl
nproc; lscpu | grep -E "Model name|NUMA|Socket" ; nvidia-smi topo -m | head -20

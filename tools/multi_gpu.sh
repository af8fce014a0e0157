#!/bin/sh
# GPU box, N GPUs (gpurun --gpus N -- sh tools/multi_gpu.sh N): bench line and batch sweep under torchrun.
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/r01_bench_${N}gpu.json
$TR tools/batch_sweep.py --batches 1024,4096,10240,16384,65536 --out gpurun_out/r01_batch_sweep_${N}gpu.jsonl > gpurun_out/sweep_${N}gpu.txt 2>&1
cut -c1-400 gpurun_out/r01_bench_${N}gpu.json; tail -14 gpurun_out/sweep_${N}gpu.txt

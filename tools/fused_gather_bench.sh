#!/bin/sh
# GPU box with N GPUs (gpurun --gpus N -- sh tools/fused_gather_bench.sh N): correctness of the fused
# solve + all-gather over NVLink and its step time next to solve-only and solve + NCCL all-gather.
N=${1:-2}
mkdir -p gpurun_out
: > gpurun_out/r01_fused_gather_${N}gpu.jsonl
for cfg in "C29 4096 f64" "C29 65536 f64" "C29 10240 f32" "C38c 8192 f64"; do
  set -- $cfg
  OSB_SHAPE=$1 OSB_BATCH=$(( $2 * N )) OSB_PRECISION=$3 OSB_STEPS=30 OSB_ORACLE=0 OSB_TOL=1 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    tests/fused_gather_worker.py 2>gpurun_out/fused_${N}gpu.err | grep '^{' | tee -a gpurun_out/r01_fused_gather_${N}gpu.jsonl
done
tail -5 gpurun_out/fused_${N}gpu.err

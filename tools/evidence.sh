#!/bin/sh
# Runs on the GPU box (gpurun -- sh tools/evidence.sh): everything profiles/ is built from.
# Numbers printed under ncu are never bench values; the bench lines come from the plain runs.
# Multi-GPU evidence: tools/multi_gpu.sh N (bench line + batch sweep under torchrun),
# tools/fused_gather_bench.sh N (solve fused with the all-gather against solve + NCCL); static SASS
# evidence needs no GPU: python tools/sass_evidence.py; quick A/B of a kernel change: tools/quick_gpu.sh.
# Afterwards, here: python tools/ncu_summary.py gpurun_out/r01_rr_final.ncu-rep profiles/r01_solve_rr
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 3 2>gpurun_out/bench_ref.err | tail -1 > gpurun_out/r01_bench_1gpu_reference.json
python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/r01_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_rr.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rr_solve -c 1 -o gpurun_out/r01_rr_final \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
python tools/batch_sweep.py --out gpurun_out/r01_batch_sweep_1gpu.jsonl > gpurun_out/sweep1.txt 2>&1
python tools/shape_sweep.py --steps 30 --seed 2026 > gpurun_out/r01_shape_sweep_auto.jsonl 2>&1
OSB_KERNEL=generic python tools/shape_sweep.py --steps 30 --seed 2026 > gpurun_out/r01_shape_sweep_generic.jsonl 2>&1

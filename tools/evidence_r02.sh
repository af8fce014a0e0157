#!/bin/sh
# Round-2 evidence pack, runs on the GPU box: gpurun --timeout 2400 -- sh tools/evidence_r02.sh
# Everything profiles/r02_* is built from. Numbers printed under ncu / compute-sanitizer are never bench values.
O=gpurun_out
mkdir -p $O
export LD_LIBRARY_PATH=$PWD/opensot_b200/lib:$PWD/oracle/_ref
set -x
# 1. the GPU test-suite (parity tests proper, through the C ABI)
python -m pytest tests -q -s -m gpu > $O/r02_pytest_gpu.txt 2>&1
# 2. bench lines: both arms on the headline config, then the other BASELINE configs
python bench.py --impl reference --steps 5 --warmup 3 2>$O/bench.err | tail -1 > $O/r02_bench_1gpu_reference.json
python bench.py 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu.json
python bench.py --shape C38c --batch 8192 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu_c38c.json
python bench.py --shape C35f 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu_c35f.json
python bench.py --dtype f32 --batch 10240 --no-cpu-baseline 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu_mpc_f32.json
python bench.py --batch 16384 --no-cpu-baseline 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu_b16384.json
# 3. launch list of the default bench command (shares, not absolutes) and of the state-input path
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > $O/launch_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_states.csv python tools/states_profile.py > $O/states_profile.log 2>&1
# 5. compute-sanitizer: every kernel family on a small batch
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > $O/r02_sanitizer_memcheck.txt 2>&1
SAN_BATCH=8 timeout 900 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_smoke.py > $O/r02_sanitizer_racecheck.txt 2>&1
# 6. sweeps and helper kernels
python tools/shape_sweep.py --steps 30 --seed 2026 > $O/r02_shape_sweep_auto.jsonl 2>&1
OSB_KERNEL=rr python tools/shape_sweep.py --steps 30 --seed 2026 --shapes C38c:8192,C35f:4096,C35h:4096 > $O/r02_shape_sweep_rr2.jsonl 2>&1
OSB_KERNEL=generic python tools/shape_sweep.py --steps 30 --seed 2026 > $O/r02_shape_sweep_generic.jsonl 2>&1
python tools/batch_sweep.py --out $O/r02_batch_sweep_1gpu.jsonl > $O/sweep_1gpu.txt 2>&1
python tools/aux_kernels_bench.py > $O/r02_aux_kernels.jsonl 2>$O/aux.err
python tools/diff_kernels.py > $O/r02_diff_kernels.txt 2>&1
# 7. the host C++ path on the same AutoStack objects
for s in C29:10 C38c:5 C35f:5; do
  ./tests/cpp/_bin/bench_dropin ${s%%:*} 4096 0.1 31 ${s##*:} 2>/dev/null | grep "^{" >> $O/r02_bench_dropin.jsonl
  ./tests/cpp/_bin/bench_dropin ${s%%:*} 4096 0.1 31 ${s##*:} 16 2>/dev/null | grep "^{" >> $O/r02_bench_dropin.jsonl
done
# 8. development sweeps quoted in DESIGN.md (occupancy, chunk pattern of the blocking call, PCIe probe)
sh tools/occupancy_sweep.sh > /dev/null 2>&1
python tools/pipe_pattern_sweep.py > $O/r02_pipe_pattern_sweep.txt 2>&1
python tools/pcie_probe_wc.py 2>&1 | grep "H2D" > $O/r02_pcie_probe_1gpu.txt

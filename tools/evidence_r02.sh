#!/bin/sh
# Round-2 evidence pack, runs on the GPU box: gpurun --timeout 1500 -- sh tools/evidence_r02.sh [tag]
# Numbers printed under ncu / compute-sanitizer are never bench values.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
set -x
# compute-sanitizer: every kernel family on a small batch (memcheck, racecheck)
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > $O/${TAG}_sanitizer_memcheck.txt 2>&1
SAN_BATCH=8 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py > $O/${TAG}_sanitizer_racecheck.txt 2>&1
# helper kernels: achieved GB/s + ncu
python tools/aux_kernels_bench.py > $O/${TAG}_aux_kernels.jsonl 2>$O/aux.err
AUX_BATCH=16384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sot_gather|sot_posture' -s 8 -c 2 -o $O/${TAG}_aux \
    python tools/aux_kernels_bench.py > $O/ncu_aux.log 2>&1
# solver kernels other than the headline instance: one full capture each
for item in C35f:4096:f64 C38c:8192:f64 C29:4096:f32 C64c:1024:f64; do
  S=$(echo $item | cut -d: -f1); B=$(echo $item | cut -d: -f2); P=$(echo $item | cut -d: -f3)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solve' -s 3 -c 1 -o $O/${TAG}_${S}_${P} \
      python tools/shape_sweep.py --shapes $S:$B --prec $P --steps 2 --seed 2026 > $O/ncu_${S}_${P}.log 2>&1
done
OSB_KERNEL=generic timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solve' -s 3 -c 1 -o $O/${TAG}_C29_generic \
    python tools/shape_sweep.py --shapes C29:4096 --prec f64 --steps 2 --seed 2026 > $O/ncu_C29_generic.log 2>&1
python tools/shape_sweep.py --steps 30 --seed 2026 > $O/${TAG}_shape_sweep_auto.jsonl 2>&1

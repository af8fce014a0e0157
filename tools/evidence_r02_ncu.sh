#!/bin/sh
# Round-2 ncu captures (gpurun brings back at most 64 MiB: run in two halves, `sh tools/evidence_r02_ncu.sh a|b`),
# then here: python tools/ncu_summary.py gpurun_out/r02_<name>.ncu-rep profiles/r02_ncu_<name> --batch B
O=gpurun_out
mkdir -p $O
set -x
if [ "$1" = "a" ]; then
for item in C29:4096:f64 C38c:8192:f64 C35f:4096:f64 C64c:1024:f64; do
  S=$(echo $item | cut -d: -f1); B=$(echo $item | cut -d: -f2); P=$(echo $item | cut -d: -f3)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solve' -s 3 -c 1 -o $O/r02_${S}_${P} \
      python tools/shape_sweep.py --shapes $S:$B --prec $P --steps 2 --seed 2026 > $O/ncu_${S}_${P}.log 2>&1
done
else
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solve' -s 3 -c 1 -o $O/r02_C29_f32 \
    python tools/shape_sweep.py --shapes C29:4096 --prec f32 --steps 2 --seed 2026 > $O/ncu_C29_f32.log 2>&1
OSB_KERNEL=generic timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solve' -s 3 -c 1 -o $O/r02_C29_generic \
    python tools/shape_sweep.py --shapes C29:4096 --prec f64 --steps 2 --seed 2026 > $O/ncu_C29_generic.log 2>&1
AUX_BATCH=16384 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sot_gather|sot_posture' -s 8 -c 2 -o $O/r02_aux \
    python tools/aux_kernels_bench.py > $O/ncu_aux.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sot_kinematics' -s 3 -c 1 -o $O/r02_kinematics \
    python tools/states_profile.py > $O/ncu_kin.log 2>&1
fi

#!/bin/sh
# Development aid: steady-state throughput of the C29 kernels against resident warps per SM (the dynamic shared memory of
# a launch is padded through OSB_DEV_SMEM_PAD until fewer CTAs fit). gpurun -- sh tools/occupancy_sweep.sh
export LD_LIBRARY_PATH=$PWD/opensot_b200/lib
O=gpurun_out/r02_occupancy_sweep.jsonl
: > $O
for pad in 0 1800 3900 6500 9700 13000 19400; do
  OSB_DEV_SMEM_PAD=$pad python tools/shape_sweep.py --shapes C29:16384 --prec f64 --steps 20 --seed 2026 >> $O
done
for pad in 0 1500 4288 6400 9140 18880; do
  OSB_DEV_SMEM_PAD=$pad python tools/shape_sweep.py --shapes C29:16384 --prec f32 --steps 20 --seed 2026 >> $O
done
cat $O

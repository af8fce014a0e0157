set -x
mkdir -p gpurun_out
for item in C35f:4096:f64 C38c:8192:f64 C64c:1024:f64; do
  S=$(echo $item | cut -d: -f1); B=$(echo $item | cut -d: -f2); P=$(echo $item | cut -d: -f3)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'solve' -s 3 -c 1 -o gpurun_out/r02b_${S}_${P}_mw \
      python tools/shape_sweep.py --shapes $S:$B --prec $P --steps 2 --seed 2026 > gpurun_out/ncu_${S}_${P}.log 2>&1
done
python -m pytest tests/test_gpu_parity.py -x -q -k "every_register_row" 2>&1 | tail -3

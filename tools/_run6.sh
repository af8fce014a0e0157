set -x
python tools/shape_sweep.py --steps 50 --seed 2026 --shapes C29:4096,C29:16384 --prec f64 2>&1 | cut -c1-300
OSB_RR16=1 python tools/shape_sweep.py --steps 50 --seed 2026 --shapes C29:4096,C29:16384 --prec f64 2>&1 | cut -c1-300

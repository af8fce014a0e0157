python tools/shape_sweep.py --steps 50 --seed 2026 --prec f64 --shapes C29:4096,C38c:8192,C35f:4096 2>&1 | cut -c1-200
python tools/states_profile.py | cut -c1-100

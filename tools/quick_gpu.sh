#!/bin/sh
# Development aid (GPU box): parity tests of the kernels, then one device-resident timing per shape.
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/shape_sweep.py --steps 30 --seed 2026 --shapes C29:4096,C38c:8192,C35f:4096,C35h:4096 2>&1 | cut -c1-200

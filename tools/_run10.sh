python -m pytest tests/test_gpu_kinematics.py -x -q -s -m gpu 2>&1 | grep "saturated\|state-input\|passed\|failed\|Error" | cut -c1-700

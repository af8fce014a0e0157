set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_live_reference.py -x -q > gpurun_out/r02_pytest_mw.txt 2>&1
tail -3 gpurun_out/r02_pytest_mw.txt
python tools/shape_sweep.py --steps 30 --seed 2026 --shapes C38c:8192,C35f:4096,C35h:4096,C64c:1024 > gpurun_out/r02_sweep_mw.jsonl 2>&1
OSB_KERNEL=rr python tools/shape_sweep.py --steps 30 --seed 2026 --shapes C38c:8192,C35f:4096,C35h:4096 > gpurun_out/r02_sweep_rr2.jsonl 2>&1
cut -c1-330 gpurun_out/r02_sweep_mw.jsonl gpurun_out/r02_sweep_rr2.jsonl
SAN_BATCH=8 timeout 900 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_smoke.py > gpurun_out/r02c_sanitizer_racecheck.txt 2>&1
grep -c "Race reported" gpurun_out/r02c_sanitizer_racecheck.txt; grep "RACECHECK SUMMARY\|Error\|error" gpurun_out/r02c_sanitizer_racecheck.txt | head

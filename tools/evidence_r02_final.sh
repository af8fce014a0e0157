#!/bin/sh
# Final evidence pack of round 2 at HEAD (short: the GPU budget of the round is nearly spent):
#   gpurun --timeout 420 -- sh tools/evidence_r02_final.sh
O=gpurun_out
mkdir -p $O
export LD_LIBRARY_PATH=$PWD/opensot_b200/lib:$PWD/oracle/_ref
timeout 200 python -m pytest tests -q -s -m gpu > $O/r02_pytest_gpu.txt 2>&1
tail -1 $O/r02_pytest_gpu.txt
timeout 120 python bench.py 2>$O/bench.err | tail -1 > $O/r02_bench_1gpu.json
timeout 60 python bench.py --shape C38c --batch 8192 --no-cpu-baseline 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu_c38c.json
timeout 60 python bench.py --shape C35f --no-cpu-baseline 2>>$O/bench.err | tail -1 > $O/r02_bench_1gpu_c35f.json
timeout 120 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > $O/r02_sanitizer_memcheck.txt 2>&1
SAN_BATCH=8 timeout 120 compute-sanitizer --tool racecheck --print-limit 100 python tools/sanitize_smoke.py > $O/r02_sanitizer_racecheck.txt 2>&1
grep -h "SUMMARY" $O/r02_sanitizer_memcheck.txt $O/r02_sanitizer_racecheck.txt
cut -c1-300 $O/r02_bench_1gpu.json

#!/bin/sh
# GPU box, N GPUs (gpurun --gpus N -- sh tools/multi_gpu_r02.sh N): both bench arms' lines and the batch sweep under torchrun.
N=${1:-4}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 20 --warmup 5 2>$O/bench_${N}gpu.err | tail -1 > $O/r02_bench_${N}gpu.json
$TR bench.py --gpus $N --steps 20 --warmup 5 --strong --batch 4096 --no-parity 2>>$O/bench_${N}gpu.err | tail -1 > $O/r02_bench_${N}gpu_strong4096.json
$TR bench.py --gpus $N --steps 20 --warmup 5 --strong --batch 10240 --dtype f32 --no-parity 2>>$O/bench_${N}gpu.err | tail -1 > $O/r02_bench_${N}gpu_mpc_f32.json
$TR bench.py --gpus $N --steps 20 --warmup 5 --shape C38c --batch 8192 --no-parity 2>>$O/bench_${N}gpu.err | tail -1 > $O/r02_bench_${N}gpu_c38c.json
$TR tools/batch_sweep.py --batches 256,1024,4096,16384,65536 --out $O/r02_batch_sweep_${N}gpu.jsonl > $O/sweep_${N}gpu.txt 2>&1
python - <<PY
import json
for f in ("r02_bench_${N}gpu.json","r02_bench_${N}gpu_strong4096.json","r02_bench_${N}gpu_mpc_f32.json","r02_bench_${N}gpu_c38c.json"):
    try:
        l=json.loads(open("$O/"+f).read().strip().splitlines()[-1])
        print(f, "value %.4g"%l["value"], "ms %.4f"%l["ms_per_step"], "e2e %.4g"%l.get("e2e",{}).get("value",0), "states", l.get("e2e_states",{}).get("value"), "fused", l.get("fused_allgather",{}).get("ms_per_step"), "nccl ag", l.get("allgather_x_ms"))
    except Exception as ex: print(f, "ERR", ex)
PY
tail -42 $O/sweep_${N}gpu.txt

#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (stack -> H,g per level -> hierarchical QP cascade)
over one batch of synthetic stacks: configs[1] of BASELINE.json, COMAN-like 29-DoF 3-level
stack (shape C29 of SURVEY.md section 8), batch = 4096 instances per GPU, fp64.
Weak scaling: every rank solves its own 4096 instances, no data-path collective.

Prints ONE JSON line (rank 0). `value` is device-timed with inputs resident in HBM; `e2e`
is the same metric through the public host-buffer entry point (osb_solve_host: pinned host
blobs -> H2D -> kernel -> D2H of x and status inside the timed region).
--impl reference times the reference's own CPU path (oracle/_ref: unmodified iHQP+qpOASES,
multi-process over the host cores) on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = "C29"
BATCH = 4096
TIGHT = 0.10          # ~3 saturated bounds per instance (SURVEY.md 6.2 first row)
EPS_FACTOR = 2e2      # iHQP default, include/OpenSoT/solvers/iHQP.h:32
SEED = 2026
LEVELS = 3
# algorithmic work per C29 stack, SURVEY.md section 8(d) (definitions restated in DESIGN.md)
FLOP_PER_STACK = 319577.0
BYTES_PER_STACK = 4528.0
METRIC = "whole-body QP solves/sec (batch x levels), 29-DoF 3-level stack"
UNIT = "QP solves/s"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_throughput(count, cycles, warmup):
    """The reference's own CPU path on this box: best of {all cores, 1 core} worker processes."""
    from oracle import portlib as P
    if not P.have_ref():
        raise RuntimeError("oracle/_ref/ref_driver missing (built by __graft_entry__.build() "
                           "where /root/reference exists; travels with the snapshot)")
    best = None
    for nproc in sorted({host_cores(), max(1, host_cores() // 2), 1}, reverse=True):
        r = P.ref_bench(SHAPE, EPS_FACTOR, SEED, count, TIGHT, nproc, cycles, warmup)
        if best is None or r["stacks_per_s"] > best["stacks_per_s"]:
            best = r
    return best


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "10"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    # bounded sample: `count` persistent solver objects, each solved once per step
    count = 512 if cores < 16 else 2048
    r = reference_throughput(count, args.steps, args.warmup)
    stacks_per_s = r["stacks_per_s"]
    value = stacks_per_s * LEVELS
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["max_busy_s"] / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C29 (n=29, levels 3/12/29I, 2 bound sets), sample of %d stacks per step, "
                               "tight=%.2f, eps_factor=%g" % (count, TIGHT, EPS_FACTOR)},
        "stacks_per_s": stacks_per_s,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["nproc"], "kind": "reference",
                         "sample": "%d stacks x %d cycles, iHQP+qpOASES hot-started, %d worker processes "
                                   "(host has %d cores); update()+solve() timed with steady_clock"
                                   % (count, args.steps, r["nproc"], cores),
                         "failures": r["failures"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import opensot_b200 as ob
    from opensot_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    desc = synth.shape(SHAPE, EPS_FACTOR)
    eng = ob.BatchedStackSolver(desc, B, device=local_rank)
    stride, n = eng.stride, eng.n

    # inputs: NSETS distinct batches (> L2 in total) rotated between steps, so every timed
    # step reads its blobs from HBM; each rank has its own instances
    blob_bytes = B * stride * 8
    nsets = max(2, int(np.ceil(160e6 / blob_bytes)))
    host_sets, dev_sets = [], []
    for s in range(nsets):
        h = torch.from_numpy(synth.fill(desc, SEED, (rank * nsets + s) * B, B, TIGHT)).pin_memory()
        host_sets.append(h)
        dev_sets.append(h.to(dev, non_blocking=True))
    d_x = torch.empty((B, n), dtype=torch.float64, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    d_it = torch.empty(B, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream(dev)

    def step(i):
        eng.solve_device(dev_sets[i % nsets], d_x, d_st, d_iters=d_it, stream=stream)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    evs[0].record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
        evs[i + 1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = evs[0].elapsed_time(evs[-1])
    kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    n_bad = int((d_st != 0).sum().item())
    mean_iters = float(d_it.double().mean().item())

    # end-to-end through the host-buffer entry point (pinned host blobs)
    x_host = torch.empty((B, n), dtype=torch.float64).pin_memory()
    st_host = torch.empty(B, dtype=torch.int32).pin_memory()
    hs = eng.host_stream()
    for i in range(3):
        eng.solve_host(host_sets[i % nsets].numpy(), x_host.numpy(), st_host.numpy())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0.record(hs)
    for i in range(args.steps):
        eng.solve_host(host_sets[(3 + i) % nsets].numpy(), x_host.numpy(), st_host.numpy())
    e1.record(hs)
    e1.synchronize()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall_ms)   # the slower of device and host clocks
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    gather_ms = None
    if world > 1:   # optional: whole-batch solution on every rank (one NCCL all-gather)
        full = torch.empty((world * B, n), dtype=torch.float64, device=dev)
        for _ in range(3):
            dist.all_gather_into_tensor(full, d_x)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier()
        g0.record()
        for _ in range(10):
            dist.all_gather_into_tensor(full, d_x)
        g1.record(); torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms = float(gt[0])

    fused = None
    if world > 1:   # the same, fused into the solve: peer stores over NVLink + device-side barrier
        from opensot_b200.distributed import FusedAllGather
        try:
            fg = FusedAllGather(eng, world * B)
            for i in range(args.warmup):
                fg.solve(dev_sets[i % nsets], stream=stream)
            torch.cuda.synchronize(); dist.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(stream)
            for i in range(args.steps):
                fg.solve(dev_sets[(args.warmup + i) % nsets], stream=stream)
            f1.record(stream); torch.cuda.synchronize()
            # the last fused step solved the same input set as the last timed step above
            differs = 0.0 if torch.equal(fg.x_all[rank * B:(rank + 1) * B], d_x) else 1.0
            ft = torch.tensor([f0.elapsed_time(f1) / args.steps, float(fg.failures()), differs], dtype=torch.float64, device=dev)
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
            fused = {"ms_per_step": float(ft[0]), "barrier_failures": int(ft[1]), "matches_unfused": ft[2].item() == 0.0,
                     "what": "solve kernel stores x/status into every rank's buffer (peer stores) + device-side barrier; "
                             "every rank holds the whole batch after each step"}
        except RuntimeError as ex:
            fused = {"unavailable": str(ex)[:300]}

    if rank == 0:
        stacks = world * B * args.steps
        stacks_per_s = stacks / (total_ms * 1e-3)
        value = stacks_per_s * LEVELS
        e2e_value = stacks / (e2e_ms * 1e-3) * LEVELS
        avg_launch_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        fp64_peak = ob.measure_fp64_tflops(local_rank, 8192)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach_tf = B * FLOP_PER_STACK / avg_launch_s / 1e12
        ach_gbs = B * BYTES_PER_STACK / avg_launch_s / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except OSError:
            pass
        info = eng.info()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C29: n=29, levels 3/12/29I (W: I, diag, I), 2 bound sets, eps_factor=%g, "
                                   "tight=%.2f, batch=%d per GPU" % (EPS_FACTOR, TIGHT, B),
                       "global_batch": world * B, "parallelism": "dp%d (batch sharded, no collective)" % world,
                       "l2": "%d input sets rotated (%.0f MB > 126 MB L2)" % (nsets, nsets * blob_bytes / 1e6)},
            "stacks_per_s": stacks_per_s, "unsolved_instances": n_bad, "mean_working_set_changes": mean_iters,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * stride * 8,
                    "d2h_bytes_per_step": B * n * 8 + B * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": ach_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": ach_tf / fp64_peak if fp64_peak > 0 else None,
                         "peak_source": "DFMA probe kernel measured in this run (MEASURED_PEAKS.json has no fp64 figure)",
                         "flop_per_stack": FLOP_PER_STACK, "kernel": eng.kernel_name(),
                         "avg_launch_ms": avg_launch_s * 1e3, "traffic": traffic},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": ach_gbs / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                             "bytes_per_stack": BYTES_PER_STACK, "traffic": traffic},
            "kernel": {"regs": info["kernel_regs"], "smem_per_cta": info["smem_per_cta"],
                       "ctas_per_sm": info["ctas_per_sm"], "grid": B, "block": 32},
        }
        if gather_ms is not None:
            line["allgather_x_ms"] = gather_ms
        if fused is not None:
            if "ms_per_step" in fused:
                fused["value"] = world * B / (fused["ms_per_step"] * 1e-3) * LEVELS
                fused["unit"] = UNIT
            line["fused_allgather"] = fused
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = host_cores()
                count = 512 if cores < 16 else 2048
                r = reference_throughput(count, 3, 1)
                line["cpu_baseline"] = {
                    "value": r["stacks_per_s"] * LEVELS, "unit": UNIT, "cores": r["nproc"], "kind": "reference",
                    "sample": "%d stacks x 3 cycles (+1 warm-up), unmodified iHQP+qpOASES (oracle/_ref), "
                              "%d worker processes of %d host cores" % (count, r["nproc"], cores)}
            except Exception as ex:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "unavailable: %s" % ex}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

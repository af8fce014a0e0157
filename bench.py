#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--shape C29|C35f|C38c|C35h] [--dtype f64|f32] [--batch B] [--strong]

A "step" is one pass of the hot path (stack -> H,g per level -> hierarchical QP cascade) over one
batch of synthetic stacks. Default = configs[1] of BASELINE.json, the configuration the metric is
quoted on: COMAN-like 29-DoF 3-level stack (shape C29 of SURVEY.md section 8), batch = 4096
instances per GPU, fp64. The other BASELINE configurations are reachable with flags:
  configs[2]  --shape C38c --batch 8192                     (38 DoF, 26 one-sided rows)
  configs[3]  --shape C29 --dtype f32 --batch 10240 --strong --gpus 4   (MPC: 20 x 512 rollouts)
  configs[4]  --batch 256..65536, --dtype f64|f32, --gpus 1|2|4|8, with and without --strong
Weak scaling (default): every rank solves its own `batch` instances. --strong: `batch` is the global
batch, cut into contiguous slices over the ranks. No data-path collective either way.

Prints ONE JSON line (rank 0). `value` is device-timed with inputs resident in HBM; `e2e` is the
same metric through the public host-buffer entry points: every step copies ITS pinned host blobs to the
device, solves, and copies ITS x and status back, all inside the timed region. `e2e.value` keeps two such
calls in flight (osb_solve_host_submit / osb_solve_wait: the complete blobs of step k+1 cross PCIe while step k is solved),
`e2e.blocking` is one blocking osb_solve_host per step, `e2e.wire_format` uploads the fields this workload never varies once
(osb_wire_configure) and sends the rest every step; `parity` compares a sample of the TIMED
batch with the live unmodified reference. --impl reference times the reference's own CPU path
(oracle/_ref: unmodified iHQP+qpOASES, multi-process over the host cores) on the same config.
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

TIGHT = 0.10          # ~3 saturated bounds per instance (SURVEY.md 6.2 first row)
EPS_FACTOR = 2e2      # iHQP default, include/OpenSoT/solvers/iHQP.h:32
SEED = 2026
UNIT = "QP solves/s"
# (n, task rows per level, identity flags, c_eq, c_in, max c_loc) of the named shapes
SHAPES = {
    "C29": dict(n=29, m=(3, 12, 29), ident=(0, 0, 1), c_eq=0, c_in=0, c_loc=0,
                what="n=29, levels 3/12/29I (W: I, diag, I), 2 bound sets"),
    "C35f": dict(n=35, m=(3, 12, 35), ident=(0, 0, 1), c_eq=12, c_in=0, c_loc=0,
                 what="n=35, levels 3/12/35I, 12 contact equalities, 2 bound sets (coman_ik 3_LEVELS shape)"),
    "C38c": dict(n=38, m=(3, 12, 38), ident=(0, 0, 1), c_eq=0, c_in=26, c_loc=0,
                 what="n=38, levels 3/12/38I, 26 one-sided inequality rows (collision + friction cones), 2 bound sets"),
    "C35h": dict(n=35, m=(6, 12, 35), ident=(0, 0, 1), c_eq=0, c_in=0, c_loc=4,
                 what="n=35, levels 6/12/35I, 4 task-local hull rows per level, 2 bound sets"),
}


def work_per_stack(shape, changes=6):
    """Algorithmic flops and bytes of one stack by the definitions of SURVEY.md section 8(d) (restated in
    DESIGN.md 4.3). Reproduces the survey's constants: C29 319 577 flop / 4 528 B, C35f 370 193 / 8 992,
    C38c 712 118 / 14 216 (tests/test_bench_contract.py)."""
    s = SHAPES[shape]
    n, flops, scalars = s["n"], 0.0, 0
    e = s["c_eq"]
    for k, m in enumerate(s["m"]):
        z = n - e
        if not s["ident"][k]:
            flops += m * n * (n + 1) + 2 * m * n            # cost assembly
            scalars += m * n
        scalars += m                                        # b
        if k > 0:
            flops += 2 * s["m"][k - 1] * n                  # optimality rhs
        flops += 2 * e * e * (n - e / 3.0)                  # QR of the equality block
        flops += 2 * n * n * z + 2 * n * z * z              # dense projected Hessian
        flops += z ** 3 / 3.0                               # Cholesky
        flops += 2 * (2 * n * z + z * z) + 2 * n * e + e * e  # solves
        flops += 2 * (s["c_in"] + s["c_loc"]) * n           # inequality evaluation
        scalars += s["c_loc"] * (n + 2)
        e += m
    flops += changes * 10 * n * n
    scalars += 2 * n + (s["c_eq"] + s["c_in"]) * (n + 2)
    return float(round(flops)), (scalars + n) * 8.0


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def metric_name(shape):
    s = SHAPES[shape]
    return "whole-body QP solves/sec (batch x levels), %d-DoF %d-level stack" % (s["n"], len(s["m"]))


def config_of(args, world):
    """The workload, identical for both arms."""
    per_rank = args.batch // world if args.strong else args.batch
    return {"workload": "%s: %s, eps_factor=%g, tight=%.2f, batch=%d per GPU"
                        % (args.shape, SHAPES[args.shape]["what"], EPS_FACTOR, TIGHT, per_rank),
            "global_batch": per_rank * world,
            "parallelism": "dp%d (batch sharded, no collective)" % world}


def reference_throughput(shape, count, cycles, warmup):
    """The reference's own CPU path on this box: best of {all cores, half, 1 core} worker processes."""
    from oracle import portlib as P
    if not P.have_ref():
        raise RuntimeError("oracle/_ref/ref_driver missing (built by __graft_entry__.build() "
                           "where /root/reference exists; travels with the snapshot)")
    best = None
    for nproc in sorted({host_cores(), max(1, host_cores() // 2), 1}, reverse=True):
        r = P.ref_bench(shape, EPS_FACTOR, SEED, count, TIGHT, nproc, cycles, warmup)
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
    levels = len(SHAPES[args.shape]["m"])
    cfg = config_of(args, world)
    count = cfg["global_batch"]          # the same stacks per step as our arm solves
    r = reference_throughput(args.shape, count, args.steps, args.warmup)
    stacks_per_s = r["stacks_per_s"]
    value = stacks_per_s * levels
    line = {
        "impl": "reference", "metric": metric_name(args.shape), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["max_busy_s"] / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "stacks_per_s": stacks_per_s,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["nproc"], "kind": "reference",
                         "sample": "%d persistent solver objects x %d cycles (+%d warm-up), iHQP+qpOASES hot-started across "
                                   "time-coherent cycles, %d worker processes (host has %d cores); update()+solve() timed with "
                                   "steady_clock" % (count, args.steps, args.warmup, r["nproc"], cores),
                         "failures": r["failures"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.dtype != "f64":
        line["note"] = "the reference computes in fp64 only; this arm is its fp64 path on the same stacks"
    print(json.dumps(line))


def parity_block(desc, layout, blobs64, x, status, dtype):
    """A sample of the TIMED batch against the live unmodified reference (oracle/_ref, iHQP+qpOASES):
    cold and hot-started; the stable set is defined by the reference alone (oracle/portlib.py
    live_stable_mask: hot == cold to 1e-8 and the answer inside its own constraints to 1e-9)."""
    import numpy as np
    from oracle import portlib as P
    okh, okc, xh, xc = P.ref_solve_both(desc, blobs64)
    stable, differ, infeas = P.live_stable_mask(desc, layout, blobs64, okh, okc, xh, xc)
    err = np.abs(x.astype(np.float64) - xh).max(axis=1)
    uns = ~stable
    out = {"sample": int(blobs64.shape[0]), "against": "oracle/_ref: unmodified iHQP+qpOASES, cold and hot-started",
           "n_stable": int(stable.sum()), "max_err_stable": float(err[stable].max()) if stable.any() else None,
           "median_err_stable": float(np.median(err[stable])) if stable.any() else None,
           "n_unstable": int(uns.sum()), "n_hot_ne_cold": int(differ.sum()), "n_ref_outside_own_constraints": int(infeas.sum()),
           "n_ref_failed": int((~(okh & okc)).sum()),
           "worst_unstable": float(err[uns & okh].max()) if (uns & okh).any() else 0.0,
           "gpu_unsolved_on_stable": int((status[stable] != 0).sum()),
           "tolerance": 1e-6 if dtype == "f64" else None}
    if uns.any():   # where the reference is not well-defined: distance to the exact optimum (restatement)
        idx = np.where(uns)[0]
        xp, stp, _ = P.port_solve(desc, blobs64[idx])
        okp = (stp == 0) & (status[idx] == 0)
        out["unstable_max_err_vs_restatement"] = float(np.abs(x[idx][okp].astype(np.float64) - xp[okp]).max(initial=0.0))
    return out


def states_e2e(eng, B, n, steps, rank, world, dist):
    """osb_solve_states_host on the COMAN tree fixture (tests/golden/robots/coman_tree.npz): pinned q in, x out."""
    import numpy as np
    import torch
    from opensot_b200 import kinematics as K
    tree = K.load_npz(os.path.join(ROOT, "tests", "golden", "robots", "coman_tree.npz"))
    if tree.n_dof != n:
        raise RuntimeError("tree / stack size mismatch")
    L = eng.layout
    eng.kinematics_create(tree)
    tmpl = np.zeros(L.stride)
    tmpl[L.w[1]:L.w[1] + 12] = [0.1] * 6 + [1.0] * 6
    pose = lambda R, p: np.concatenate([np.asarray(R).reshape(-1), p])   # noqa: E731
    # references = the poses of the home configuration (q = 0), states scattered around it so that ~2-3 bounds
    # saturate per instance, the regime of the synthetic C29 inputs (tight = 0.10)
    tasks = [dict(kind="com", level=0, lam=0.1, desired=np.array([-0.0006077468177577, -0.00012074685368631, 0.02043438727581192])),
             dict(kind="cartesian", distal="l_wrist", level=1, row0=0, lam=0.1, desired=pose(np.eye(3), [0.005305, 0.1558, -0.05035101])),
             dict(kind="cartesian", distal="r_wrist", base="Waist", level=1, row0=6, lam=0.1,
                  desired=pose(np.eye(3), [0.005305, -0.1558, -0.05035101]))]
    eng.states_configure(tmpl, tasks, q_ref=np.zeros(n), q_min=tree.q_min, q_max=tree.q_max, qdot_max=np.minimum(tree.qdot_max, 4.0),
                         lam=0.05, dT=0.01, postural_level=2, joint_limit_set=0, velocity_limit_set=1)
    rng = np.random.default_rng(SEED + rank)
    qs = [torch.from_numpy(0.12 * rng.uniform(tree.q_min, tree.q_max, size=(B, n))).pin_memory() for _ in range(4)]
    x = torch.empty((B, n), dtype=torch.float64).pin_memory()
    st = torch.empty(B, dtype=torch.int32).pin_memory()
    for i in range(3):
        eng.solve_states_host(qs[i % 4].numpy(), x.numpy(), st.numpy())
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        eng.solve_states_host(qs[i % 4].numpy(), x.numpy(), st.numpy())
    ms_blocking = (time.perf_counter() - t0) * 1e3 / steps
    x_ref = x.numpy().copy()                      # answer of the last blocking step (input set (steps - 1) % 4)
    # two calls in flight (osb_solve_states_submit / osb_solve_wait): step k+1's joint positions cross PCIe and its
    # producers start while step k is being solved; every step still copies its inputs in and its solutions out
    xs = [torch.empty((B, n), dtype=torch.float64).pin_memory() for _ in range(2)]
    sts = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    def run(count):
        for i in range(count):
            if eng.in_flight() == 2:
                eng.wait()
            eng.submit_states(qs[i % 4].numpy(), xs[i & 1].numpy(), sts[i & 1].numpy())
        while eng.in_flight():
            eng.wait()
    run(4)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    run(steps)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    same = bool(np.array_equal(xs[(steps - 1) & 1].numpy(), x_ref))
    if dist is not None:
        t = torch.tensor([ms, ms_blocking], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_blocking = float(t[0]), float(t[1])
    return {"ms_per_step": ms, "h2d_bytes_per_step": B * n * 8, "d2h_bytes_per_step": B * n * 8 + B * 4,
            "unsolved_instances": int((st != 0).sum()), "kernels_per_step": 3,
            "mode": "two calls in flight (osb_solve_states_submit / osb_solve_wait)", "blocking_ms_per_step": ms_blocking,
            "pipelined_equals_blocking": same,
            "what": "H2D of joint positions only -> sot_posture_kernel + sot_kinematics_kernel (COMAN: CoM, "
                    "l_wrist, r_wrist w.r.t. Waist, posture, joint + velocity limits) -> solver kernel -> D2H; host wall clock "
                    "over all steps including the last wait; blocking_ms_per_step is osb_solve_states_host, one call at a time"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="C29", choices=sorted(SHAPES))
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU (global batch with --strong)")
    ap.add_argument("--strong", action="store_true", help="fixed global batch, sliced over the ranks")
    ap.add_argument("--parity-sample", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.strong and args.batch % world:
        raise SystemExit("--strong needs a batch divisible by the number of ranks")

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import opensot_b200 as ob
    from opensot_b200 import synth, _abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    f64 = args.dtype == "f64"
    B = args.batch // world if args.strong else args.batch
    levels = len(SHAPES[args.shape]["m"])
    flop_per_stack, bytes_per_stack = work_per_stack(args.shape)
    if not f64:
        bytes_per_stack /= 2
    desc = synth.shape(args.shape, EPS_FACTOR if f64 else max(EPS_FACTOR, _abi.F32_EPS_FACTOR))
    eng = ob.BatchedStackSolver(desc, B, device=local_rank, precision=args.dtype)
    stride, n = eng.stride, eng.n
    es = 8 if f64 else 4
    tdt = torch.float64 if f64 else torch.float32

    # inputs: NSETS distinct batches (> L2 in total) rotated between steps, so every timed
    # step reads its blobs from HBM; each rank has its own instances
    blob_bytes = B * stride * es
    nsets = max(2, int(np.ceil(160e6 / blob_bytes)))
    nsets = min(nsets, 64)
    host_sets, dev_sets = [], []
    first_of = lambda s: (rank * nsets + s) * B   # noqa: E731
    for s in range(nsets):
        h = torch.from_numpy(eng.pack(synth.fill(desc, SEED, first_of(s), B, TIGHT))).pin_memory()
        host_sets.append(h)
        dev_sets.append(h.to(dev, non_blocking=True))
    d_x = torch.empty((B, n), dtype=tdt, device=dev)
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
    last_set = (args.warmup + args.steps - 1) % nsets       # the inputs of the last timed step
    x_last = d_x.cpu().numpy()
    st_last = d_st.cpu().numpy()

    # the same K steps with consecutive steps on two streams (independent batches: the tail of one launch, where the SMs
    # run out of stacks, overlaps the head of the next); reported beside `value`, which stays one launch after the other
    s2 = torch.cuda.Stream(device=dev)
    outs = [(d_x, d_st, d_it), (torch.empty_like(d_x), torch.empty_like(d_st), torch.empty_like(d_it))]
    two = [stream, s2]
    torch.cuda.synchronize()
    oa, ob0, ob1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for rep in range(2):            # first pass warms the second stream up
        if world > 1:
            dist.barrier()
        oa.record(stream); s2.wait_event(oa)
        for i in range(args.steps):
            eng.solve_device(dev_sets[(args.warmup + i) % nsets], outs[i & 1][0], outs[i & 1][1], d_iters=outs[i & 1][2], stream=two[i & 1])
        ob0.record(stream); ob1.record(s2)
        torch.cuda.synchronize()
    overlapped_ms = max(oa.elapsed_time(ob0), oa.elapsed_time(ob1))
    eng.solve_device(dev_sets[last_set], d_x, d_st, d_iters=d_it, stream=stream)   # d_x holds the last timed step's answer again
    torch.cuda.synchronize()

    # end-to-end through the host-buffer entry point (pinned host blobs)
    e2e_ms = float("nan")
    if not args.no_e2e:
        x_host = torch.empty((B, n), dtype=tdt).pin_memory()
        st_host = torch.empty(B, dtype=torch.int32).pin_memory()
        hs = eng.host_stream()
        # warm-up: every page-locked input set crosses PCIe a few times first (measured: the first DMA passes over a freshly
        # pinned buffer run at half the link rate)
        for i in range(max(3, 2 * nsets)):
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
        e2e_blocking_ms = max(e0.elapsed_time(e1), e2e_wall_ms)   # the slower of device and host clocks
        x_blocking = x_host.numpy().copy()            # answer to input set (3 + steps - 1) % nsets
        # the same with two calls in flight (osb_solve_host_submit / osb_solve_wait): the blobs of step k+1 cross PCIe
        # while step k is solved and its solutions travel back; every step still copies its own blobs in and its own
        # solutions out, all of it inside the timed region (host wall clock up to the last wait)
        xh = [torch.empty((B, n), dtype=tdt).pin_memory() for _ in range(2)]
        sh = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
        def run_pipelined(count):
            for i in range(count):
                if eng.in_flight() == 2:
                    eng.wait()
                eng.submit_host(host_sets[(3 + i) % nsets].numpy(), xh[i & 1].numpy(), sh[i & 1].numpy())
            while eng.in_flight():
                eng.wait()
        run_pipelined(4)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        run_pipelined(args.steps)
        e2e_full_ms = (time.perf_counter() - t0) * 1e3
        e2e_same = bool(np.array_equal(xh[(args.steps - 1) & 1].numpy(), x_blocking))
        # the same with the wire format (osb_wire_configure): the fields the generator never varies -- the velocity-limit box,
        # the task weights, zero right-hand sides of contact rows, -inf sides of one-sided rows: what a robot stack sets once --
        # are uploaded once, every step sends the remaining runs of ITS blobs, a scatter kernel puts them in place
        static_segs = synth.static_segments(desc, eng.layout)
        n_static = sum(c for _, c in static_segs)
        wire_stride = stride
        try:
            eng.wire_configure(static_segs, synth.fill(desc, SEED, first_of(0), 1, TIGHT)[0])
            wire_stride = eng.wire_stride()
            # (page-locked straight from cudaHostAlloc: buffers pinned by torch's pin_memory() need several DMA passes before
            # they cross at the link rate, which made this figure depend on the warm-up)
            wire_sets = []
            for h in host_sets:
                w = ob.host_empty((B, wire_stride), h.numpy().dtype)
                eng.wire_pack(h.numpy(), out=w)
                wire_sets.append(w)
        except RuntimeError as ex:      # reported, never required (every rank sees the same shape: all or none)
            wire_sets = None
            sys.stderr.write("wire format unavailable: %s\n" % ex)
        def run_wire(count):
            for i in range(count):
                if eng.in_flight() == 2:
                    eng.wait()
                eng.submit_host_wire(wire_sets[(3 + i) % nsets], xh[i & 1].numpy(), sh[i & 1].numpy())
            while eng.in_flight():
                eng.wait()
        if wire_sets is not None:
            run_wire(max(8, 8 * nsets))   # (fresh page-locked buffers, just written by the CPU: see the warm-up note above)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            run_wire(args.steps)
            e2e_ms = (time.perf_counter() - t0) * 1e3
            e2e_same = e2e_same and bool(np.array_equal(xh[(args.steps - 1) & 1].numpy(), x_blocking))
    # second end-to-end figure, state-input path (osb_solve_states_host): the same 29-DoF 3-level stack with robot
    # meaning (COMAN: CoM / two wrists / posture, joint + velocity limits), every A, b and bound produced on the device
    # from joint positions; only q [B, n] crosses PCIe
    states = None
    if args.shape == "C29" and f64 and not args.no_e2e:
        try:
            states = states_e2e(eng, B, n, args.steps, rank, world, dist if world > 1 else None)
        except Exception as ex:  # reported, never required
            states = {"unavailable": str(ex)[:200]}
    clocks = sampler.stop() if rank == 0 else None

    if args.no_e2e:
        e2e_blocking_ms, e2e_full_ms, e2e_same = float("nan"), float("nan"), None
    t = torch.tensor([total_ms, e2e_ms, e2e_blocking_ms, overlapped_ms, e2e_full_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_blocking_ms, overlapped_ms, e2e_full_ms = (float(v) for v in t)

    gather_ms = None
    if world > 1:   # optional: whole-batch solution on every rank (one NCCL all-gather)
        full = torch.empty((world * B, n), dtype=tdt, device=dev)
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
        value = stacks_per_s * levels
        avg_launch_s = (sum(kernel_ms) / len(kernel_ms)) * 1e-3
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        if f64:
            peak_tf = ob.measure_fp64_tflops(local_rank, 8192)
            peak_src = "DFMA probe kernel measured in this run (MEASURED_PEAKS.json has no fp64 figure)"
        else:
            peak_tf = ob.measure_fp32_tflops(local_rank, 8192)
            peak_src = "FFMA probe kernel measured in this run (MEASURED_PEAKS.json has no fp32 figure)"
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        ach_tf = B * flop_per_stack / avg_launch_s / 1e12
        ach_gbs = B * bytes_per_stack / avg_launch_s / 1e9
        # DRAM traffic of this kernel per launch: from the committed `ncu --set full` capture of the
        # SAME kernel, shape, precision and batch (profiles/ncu_summary.json lists them); null otherwise
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as f:
                caps = json.load(f)
            key = "%s:%s:%d" % (args.shape, args.dtype, B)
            if key in caps.get("captures", {}):
                traffic = caps["captures"][key]["dram_bytes_per_launch"]
                traffic_src = caps["captures"][key]["source"]
        except (OSError, ValueError):
            pass
        info = eng.info()
        cfg = config_of(args, world)
        cfg["l2"] = "%d input sets rotated (%.0f MB > 126 MB L2)" % (nsets, nsets * blob_bytes / 1e6)
        line = {
            "metric": metric_name(args.shape), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "stacks_per_s": stacks_per_s, "unsolved_instances": n_bad, "mean_working_set_changes": mean_iters,
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "fp64" if f64 else "fp32", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": ach_tf / peak_tf if peak_tf > 0 else None,
                         "peak_source": peak_src,
                         "flop_per_stack": flop_per_stack, "kernel": eng.kernel_name(),
                         "avg_launch_ms": avg_launch_s * 1e3, "traffic": traffic, "traffic_source": traffic_src},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": ach_gbs / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                             "bytes_per_stack": bytes_per_stack, "traffic": traffic},
            "kernel": {"regs": info["kernel_regs"], "smem_per_cta": info["smem_per_cta"],
                       "ctas_per_sm": info["ctas_per_sm"], "grid": B, "block": 32 * info["warps_per_cta"]},
        }
        line["two_streams"] = {"value": stacks / (overlapped_ms * 1e-3) * levels, "unit": UNIT, "ms_per_step": overlapped_ms / args.steps,
                               "what": "the same K device-resident steps, consecutive steps on two CUDA streams (the tail of one "
                                       "launch overlaps the head of the next); `value` and `roofline` are one launch after the other"}
        if not args.no_e2e:
            line["e2e"] = {"value": stacks / (e2e_full_ms * 1e-3) * levels, "unit": UNIT, "h2d_bytes_per_step": B * stride * es,
                           "d2h_bytes_per_step": B * n * es + B * 4, "ms_per_step": e2e_full_ms / args.steps,
                           "mode": "two calls in flight (osb_solve_host_submit / osb_solve_wait): pinned host blobs of THIS step -> H2D "
                                   "-> solve -> D2H of x and status, every step, complete blobs; host wall clock over all steps "
                                   "including the last wait",
                           "pipelined_equals_blocking": e2e_same,
                           "wire_format": {"value": (stacks / (e2e_ms * 1e-3) * levels) if e2e_ms == e2e_ms else None,
                                           "ms_per_step": (e2e_ms / args.steps) if e2e_ms == e2e_ms else None,
                                           "h2d_bytes_per_step": B * wire_stride * es,
                                           "what": "the same with osb_wire_configure / osb_solve_host_wire_submit: %d of %d elements per "
                                                   "blob never vary in this workload (velocity-limit box, task weights, zero / -inf sides "
                                                   "of constraint rows: set once on a robot stack) and are uploaded once; every step sends "
                                                   "the other runs of ITS blobs and a scatter kernel puts them in place" % (n_static, stride)},
                           "blocking": {"value": stacks / (e2e_blocking_ms * 1e-3) * levels, "ms_per_step": e2e_blocking_ms / args.steps,
                                        "what": "osb_solve_host, one blocking call per step (the call a control loop makes)"}}
        if states is not None:
            if "ms_per_step" in states:
                states["value"] = world * B / (states["ms_per_step"] * 1e-3) * levels
                states["unit"] = UNIT
            line["e2e_states"] = states
        if not args.no_parity:
            try:
                ns = min(args.parity_sample, B)
                b64 = synth.fill(desc, SEED, first_of(last_set), ns, TIGHT)
                line["parity"] = parity_block(desc, eng.layout, b64, x_last[:ns], st_last[:ns], args.dtype)
                if not f64:
                    line["parity"]["note"] = ("fp32 engine (eps raised to %.1e) against the fp64 reference at the same eps factor"
                                              % desc.eps)
            except Exception as ex:  # the checker is reported, never required
                line["parity"] = {"unavailable": str(ex)[:200]}
        if gather_ms is not None:
            line["allgather_x_ms"] = gather_ms
        if fused is not None:
            if "ms_per_step" in fused:
                fused["value"] = world * B / (fused["ms_per_step"] * 1e-3) * levels
                fused["unit"] = UNIT
            line["fused_allgather"] = fused
        if world == 1 and not args.no_cpu_baseline:
            try:
                cores = host_cores()
                count = min(B, 512 if cores < 16 else 2048)
                r = reference_throughput(args.shape, count, 3, 1)
                line["cpu_baseline"] = {
                    "value": r["stacks_per_s"] * levels, "unit": UNIT, "cores": r["nproc"], "kind": "reference",
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

"""Parity tests proper: the CUDA path, called through the C ABI (ctypes), against
 (1) the committed golden vectors of the UNMODIFIED reference (iHQP + qpOASES),
 (2) the reference's own known-answer tests,
 (3) the CPU restatement (oracle/ihqp_port.c) on fresh seeded inputs.
Tolerance: joint-velocity infinity-norm <= 1e-6 (BASELINE.json north_star), fp64."""
import numpy as np
import pytest

import common
from opensot_b200 import _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob(native_lib):
    import opensot_b200
    if opensot_b200.lib().osb_device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box (no CPU fallback)")
    return opensot_b200


@pytest.mark.parametrize("tag", common.golden_tags())
def test_golden_reference_vectors(ob, port, tag):
    d, z = common.load_golden(tag)
    blobs = z["blobs"]
    eng = ob.BatchedStackSolver(d, blobs.shape[0], store_levels=True)
    x, st, xl = eng.solve_host(blobs)
    assert (st == 0).all(), "status words: %s" % st[st != 0]
    x_port, st_port, _ = port.port_solve(d, blobs)
    stable = common.stable_mask(z, x_port)
    err = np.abs(x - z["x_qpoases"]).max(axis=1)
    print("%s: %d/%d reference-stable, max err on stable %.2e, unstable: %s"
          % (tag, stable.sum(), len(stable), err[stable].max(), err[~stable]))
    assert stable.sum() >= 0.9 * len(stable)
    assert err[stable].max() <= common.PARITY_TOL
    # instances where the reference back-ends disagree: we must side with the exact optimum
    assert np.abs(x - x_port).max() <= common.PARITY_TOL
    L = ob.blob_layout(d)
    assert common.max_violation(d, L, blobs, x).max() <= 1e-9


def test_reference_known_answers(ob):
    for k in common.load_kats():
        d, blob = common.kat_to_stack(k, ob.blob_layout)
        eng = ob.BatchedStackSolver(d, 1)
        x, st = eng.solve_host(blob[None, :])
        assert st[0] == 0, k["name"]
        assert np.abs(x[0] - np.asarray(k["expect"])).max() <= k["tol"], (k["name"], x[0])
        assert np.abs(x[0] - np.asarray(k["x_reference"])).max() <= 1e-6, (k["name"], x[0])


@pytest.mark.parametrize("shape,tight,count", [("C29", 0.05, 512), ("C29", 0.35, 512), ("C35f", 0.2, 256),
                                               ("C38c", 0.2, 256), ("C35h", 0.2, 256), ("C64c", 0.1, 64),
                                               ("C12", 0.5, 512), ("C7c", 0.5, 512)])
def test_against_restatement(ob, port, shape, tight, count):
    d = port.shape(shape)
    blobs = port.synth(d, 1234, 0, count, tight)
    eng = ob.BatchedStackSolver(d, count, store_levels=True)
    x, st, xl = eng.solve_host(blobs)
    xp, xlp, stp, itp = port.port_solve(d, blobs, levels=True)
    assert (st == 0).all() and (stp == 0).all()
    err = np.abs(x - xp).max(axis=1)
    print("%s tight=%.2f: max err %.2e median %.2e" % (shape, tight, err.max(), np.median(err)))
    assert err.max() <= common.PARITY_TOL


def test_active_levels_match_restatement(ob, port):
    d = port.shape("C29")
    blobs = port.synth(d, 77, 0, 128, 0.2)
    eng = ob.BatchedStackSolver(d, 128)
    for active in ([1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 0, 0], [1, 1, 1]):
        eng.set_active_levels(active)
        x, st = eng.solve_host(blobs)
        xp, stp, _ = port.port_solve(d, blobs, active=active)
        assert (st == 0).all()
        assert np.abs(x - xp).max() <= common.PARITY_TOL, active


@pytest.mark.parametrize("shape,tight", [("C29w", 0.2), ("C29r", 0.2), ("C29c", 0.2), ("C20d", 0.3), ("C16e", 0.3)])
def test_every_level_path_against_reference(ob, port, shape, tight):
    """dense W, user regularisation task, linear cost terms, weighted identity levels, global
    equalities + inequalities + task-local rows: the DENSE / LQ / TRIVIAL level paths. The live
    reference (oracle/_ref, when it travelled) is the arbiter, else the restatement."""
    d = port.shape(shape)
    blobs = port.synth(d, 555, 0, 160, tight)
    eng = ob.BatchedStackSolver(d, 160)
    x, st = eng.solve_host(blobs)
    xp, stp, _ = port.port_solve(d, blobs)
    if port.have_ref():
        ok, xr, _ = port.ref_solve(d, blobs)
    else:
        ok, xr = (stp == 0), xp
    # genuinely infeasible instances (a level's local rows can contradict earlier levels'
    # optimality rows): the reference returns false, we must flag the same instances
    assert np.array_equal(st != 0, ~ok), (np.where(st != 0)[0], np.where(~ok)[0])
    assert all((s & 0xff) == _abi.INFEASIBLE for s in st[st != 0])
    good = (st == 0)
    # reference-stable set: qpOASES and the restatement agree to 1e-8;
    # on the rest (qpOASES stops ~1e-7 outside the constraints) we must match the exact optimum
    stable = good & (stp == 0) & (np.abs(xr - xp).max(axis=1) <= common.STABLE_TOL)
    err = np.abs(x - xr).max(axis=1)
    errp = np.abs(x - xp).max(axis=1)
    print("%s: %d infeasible (reference agrees), %d/%d reference-stable, max err %.2e; vs restatement %.2e"
          % (shape, (~good).sum(), stable.sum(), good.sum(), err[stable].max(), errp[good & (stp == 0)].max()))
    assert err[stable].max() <= common.PARITY_TOL
    assert errp[good & (stp == 0)].max() <= common.PARITY_TOL
    if stable.sum() < 0.95 * good.sum():
        # C29c: a linear cost term on a rank-deficient level is LP-like in that level's null
        # space (only eps holds it), so the two CPU arbiters themselves differ by ~1e-8..1e-7;
        # then every instance must still be within the parity bound of the reference itself
        assert err[good].max() <= common.PARITY_TOL


def test_edge_cases(ob, port):
    rng = np.random.default_rng(3)
    n = 6
    # inactive task (A = 0, b unchanged; Task.h:382-386): rank 0 level, next level unaffected
    d = _abi.make_desc(n, [dict(m=2), dict(m=n, kind=_abi.TASK_IDENTITY)], n_bound_sets=1)
    L = ob.blob_layout(d)
    blob = np.zeros(L.stride)
    blob[L.b[0]:L.b[0] + 2] = [1.0, -1.0]
    b1 = rng.uniform(-0.5, 0.5, n)
    blob[L.b[1]:L.b[1] + n] = b1
    blob[L.l[0]:L.l[0] + n] = -0.3; blob[L.u[0]:L.u[0] + n] = 0.3
    x, st = ob.BatchedStackSolver(d, 1).solve_host(blob[None])
    assert st[0] == 0 and np.allclose(x[0], np.clip(b1, -0.3, 0.3), atol=1e-9)
    # no bounds, no constraints: least-squares cascade = pseudo-inverse recursion
    d = _abi.make_desc(n, [dict(m=2), dict(m=n, kind=_abi.TASK_IDENTITY)], n_bound_sets=0, eps_factor=1.0)
    L = ob.blob_layout(d)
    blob = np.zeros(L.stride)
    A0 = rng.standard_normal((2, n)); b0 = rng.standard_normal(2); b1 = rng.standard_normal(n)
    blob[L.A[0]:L.A[0] + 2 * n] = A0.ravel(); blob[L.b[0]:L.b[0] + 2] = b0; blob[L.b[1]:L.b[1] + n] = b1
    x, st = ob.BatchedStackSolver(d, 1).solve_host(blob[None])
    P0 = np.linalg.pinv(A0)
    assert st[0] == 0 and np.allclose(x[0], P0 @ b0 + (np.eye(n) - P0 @ A0) @ b1, atol=1e-9)
    # duplicated equality rows are fine, contradictory ones are reported as infeasible
    d = _abi.make_desc(n, [dict(m=n, kind=_abi.TASK_IDENTITY)], c_eq=3, n_bound_sets=0)
    L = ob.blob_layout(d)
    blob = np.zeros(L.stride)
    row = rng.standard_normal(n)
    Ceq = np.vstack([row, 2 * row, rng.standard_normal(n)])
    blob[L.Ceq:L.Ceq + 3 * n] = Ceq.ravel(); blob[L.deq:L.deq + 3] = [1.0, 2.0, 0.5]
    eng = ob.BatchedStackSolver(d, 1)
    x, st = eng.solve_host(blob[None])
    assert st[0] == 0 and np.allclose(Ceq @ x[0], [1.0, 2.0, 0.5], atol=1e-9)
    blob[L.deq + 1] = 3.0
    x, st = eng.solve_host(blob[None])
    assert (st[0] & 0xff) == _abi.INFEASIBLE
    # contradictory inequality: x0 <= -1 (bound) and 1 <= x0 <= 2 (row)
    d = _abi.make_desc(n, [dict(m=n, kind=_abi.TASK_IDENTITY)], c_in=1, n_bound_sets=1)
    L = ob.blob_layout(d)
    blob = np.zeros(L.stride)
    blob[L.Cin] = 1.0; blob[L.lA] = 1.0; blob[L.uA] = 2.0
    blob[L.l[0]:L.l[0] + n] = -2.0; blob[L.u[0]:L.u[0] + n] = 2.0; blob[L.u[0]] = -1.0
    x, st = ob.BatchedStackSolver(d, 1).solve_host(blob[None])
    assert (st[0] & 0xff) == _abi.INFEASIBLE
    # infinite bounds (+-inf and +-DBL_MAX sentinels, CollisionAvoidance.cpp:104-107) are clipped
    d = _abi.make_desc(n, [dict(m=n, kind=_abi.TASK_IDENTITY)], c_in=2, n_bound_sets=1)
    L = ob.blob_layout(d)
    blob = np.zeros(L.stride)
    blob[L.b[0]:L.b[0] + n] = b1
    blob[L.lA:L.lA + 2] = [-np.inf, -np.finfo(float).max]; blob[L.uA:L.uA + 2] = [np.inf, np.finfo(float).max]
    blob[L.l[0]:L.l[0] + n] = -np.inf; blob[L.u[0]:L.u[0] + n] = np.inf
    x, st = ob.BatchedStackSolver(d, 1).solve_host(blob[None])
    assert st[0] == 0 and np.allclose(x[0], b1 / (1 + d.eps), atol=1e-12)


def test_nan_input_is_flagged_per_instance(ob, port):
    d = port.shape("C29")
    blobs = port.synth(d, 3, 0, 16, 0.1)
    L = ob.blob_layout(d)
    blobs[5, L.A[1] + 7] = np.nan
    blobs[9, L.b[2] + 2] = np.nan
    x, st = ob.BatchedStackSolver(d, 16).solve_host(blobs)
    assert st[5] == _abi.NAN_INPUT and st[9] == _abi.NAN_INPUT
    ok = np.ones(16, bool); ok[[5, 9]] = False
    assert (st[ok] == 0).all() and np.isfinite(x).all()
    xp, _, _ = port.port_solve(d, np.nan_to_num(blobs))
    assert np.abs(x[ok] - xp[ok]).max() <= common.PARITY_TOL


def test_full_size_batch_properties(ob, port):
    """BASELINE configs at full size (C29 x 4096, C38c x 8192): size-independent properties.
    feasibility, lexicographic optimality of level 0 (no feasible point does better on the
    first task), idempotence (same inputs -> bit-identical outputs), batch-order invariance."""
    for shape, B in (("C29", 4096), ("C38c", 8192)):
        d = port.shape(shape)
        L = ob.blob_layout(d)
        blobs = port.synth(d, 9, 0, B, 0.15)
        eng = ob.BatchedStackSolver(d, B, store_levels=True)
        x, st, xl = eng.solve_host(blobs)
        assert (st == 0).all()
        assert common.max_violation(d, L, blobs, x).max() <= 1e-9
        x2, st2, _ = eng.solve_host(blobs)
        assert np.array_equal(x, x2)                      # deterministic / idempotent
        perm = np.random.default_rng(0).permutation(B)
        x3, st3, _ = eng.solve_host(blobs[perm])
        assert np.array_equal(x3, x[perm])                # instances are independent
        # later levels keep the first task's value: A0 x_final == A0 x_level0
        n, m0 = d.n, d.m[0]
        A0 = blobs[:, L.A[0]:L.A[0] + m0 * n].reshape(B, m0, n)
        r_first = np.einsum("bij,bj->bi", A0, xl[:, 0])
        r_last = np.einsum("bij,bj->bi", A0, x)
        assert np.abs(r_first - r_last).max() <= 1e-9
        # a sample is checked against the restatement as well
        xs, sts, _ = port.port_solve(d, blobs[:128])
        assert np.abs(xs - x[:128]).max() <= common.PARITY_TOL


def test_device_side_piling(ob, port):
    """osb_stack_gather_device == AutoStack::update() piling: child task blocks living in
    separate [B, rows, cols] tensors land at their row offsets, with operator% row selection
    and the SubTask lambda scale (SubTask.cpp:57-58)."""
    import torch
    d = port.shape("C29")
    L = ob.blob_layout(d)
    B, n = 64, d.n
    blobs = port.synth(d, 21, 0, B, 0.1)
    ref = blobs.copy()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    # level 1 = two child tasks of 6 rows; the second one is rows [4,1,0,5,2,3] of an 8-row task, scaled by 0.5
    t_a = torch.randn((B, 6, n), dtype=torch.float64, generator=g)
    t_b = torch.randn((B, 8, n), dtype=torch.float64, generator=g)
    b_a = torch.randn((B, 6), dtype=torch.float64, generator=g)
    b_b = torch.randn((B, 8), dtype=torch.float64, generator=g)
    sel = torch.tensor([4, 1, 0, 5, 2, 3], dtype=torch.int32)
    ref[:, L.A[1]:L.A[1] + 6 * n] = t_a.reshape(B, -1).numpy()
    ref[:, L.A[1] + 6 * n:L.A[1] + 12 * n] = t_b[:, sel.long()].reshape(B, -1).numpy()
    ref[:, L.b[1]:L.b[1] + 6] = b_a.numpy()
    ref[:, L.b[1] + 6:L.b[1] + 12] = 0.5 * b_b[:, sel.long()].numpy()
    eng = ob.BatchedStackSolver(d, B)
    d_blobs = torch.from_numpy(blobs).to(dev)
    sel_d = sel.to(dev)
    srcs = [
        dict(src=t_a.to(dev), rows=6, cols=n, dst_off=L.A[1]),
        dict(src=t_b.to(dev), rows=6, cols=n, dst_off=L.A[1] + 6 * n, row_index=sel_d),
        dict(src=b_a.to(dev).reshape(B, 6, 1), rows=6, cols=1, dst_off=L.b[1]),
        dict(src=b_b.to(dev).reshape(B, 8, 1), rows=6, cols=1, dst_off=L.b[1] + 6, row_index=sel_d, scale=0.5),
    ]
    eng.gather_device(srcs, d_blobs)
    torch.cuda.synchronize()
    assert np.array_equal(d_blobs.cpu().numpy(), ref)
    d_x = torch.empty((B, n), dtype=torch.float64, device=dev)
    d_st = torch.empty(B, dtype=torch.int32, device=dev)
    eng.solve_device(d_blobs, d_x, d_st)
    torch.cuda.synchronize()
    xp, stp, _ = port.port_solve(d, ref)
    assert (d_st.cpu().numpy() == 0).all()
    assert np.abs(d_x.cpu().numpy() - xp).max() <= common.PARITY_TOL


def _f32_errors(ob, port, shape, tight, B, seed):
    # fp32 engines use eps >= OSB_F32_MIN_EPS; the oracle is given the same eps
    d = port.shape(shape, _abi.F32_EPS_FACTOR)
    blobs = port.synth(d, seed, 0, B, tight)
    eng = ob.BatchedStackSolver(d, B, precision="f32")
    b32 = eng.pack(blobs)
    x, st = eng.solve_host(b32)
    assert x.dtype == np.float32
    # the fp64 oracle sees exactly the float-rounded inputs the kernel saw
    xp, stp, _ = port.port_solve(d, b32[:, :blobs.shape[1]].astype(np.float64))
    return x.astype(np.float64), st, xp, stp


def test_joint_state_producers(ob, port):
    """Postural error, JointLimits and VelocityLimits boxes computed on device from
    joint states are bit-identical to the reference formulas restated in numpy
    (Postural.cpp:97-100 b = v_ref + lambda (q_ref - q); JointLimits.cpp:37-59
    u = max(0, (q_max - dq) s), l = min(0, (q_min - dq) s), dq = q - q_neutral;
    VelocityLimits.cpp:72-89 u = |qdot_max| dT, l = -u), and the stack solved from those blobs
    equals the stack solved from host-built blobs."""
    import torch
    d = port.shape("C29")
    L = ob.blob_layout(d)
    B, n = 512, d.n
    rng = np.random.default_rng(3)
    blobs = port.synth(d, 31, 0, B, 0.1)
    q = rng.uniform(-1.0, 1.0, (B, n)); q_ref = rng.uniform(-1.0, 1.0, n); v_ref = 0.01 * rng.standard_normal((B, n))
    q_neutral = 0.1 * rng.standard_normal(n)
    q_min = -1.5 - rng.uniform(0, 0.5, n); q_max = 1.5 + rng.uniform(0, 0.5, n); qdot = rng.uniform(-2.0, 2.0, n)
    q[:, 3] = q_max[3] + q_neutral[3] + 0.2          # beyond the limit: the box must clamp to 0
    lam, scal, dT = 0.7, 0.5, 0.004
    ref = blobs.copy()
    ref[:, L.b[2]:L.b[2] + n] = v_ref + lam * (q_ref[None, :] - q)
    dq = q - q_neutral[None, :]
    ref[:, L.u[0]:L.u[0] + n] = np.maximum((q_max[None, :] - dq) * scal, 0.0)
    ref[:, L.l[0]:L.l[0] + n] = np.minimum((q_min[None, :] - dq) * scal, 0.0)
    ref[:, L.u[1]:L.u[1] + n] = 1.0 * np.abs(qdot)[None, :] * dT
    ref[:, L.l[1]:L.l[1] + n] = -1.0 * np.abs(qdot)[None, :] * dT
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    eng = ob.BatchedStackSolver(d, B)
    d_blobs = t(blobs)
    eng.produce_posture_device(d_blobs, t(q), q_ref=t(q_ref), v_ref=t(v_ref), q_min=t(q_min), q_max=t(q_max),
                               q_neutral=t(q_neutral), qdot_max=t(qdot), lam=lam, bound_scaling=scal, dT=dT,
                               postural_level=2, joint_limit_set=0, velocity_limit_set=1)
    torch.cuda.synchronize()
    got = d_blobs.cpu().numpy()
    assert np.array_equal(got, ref)
    assert (got[:, L.u[0] + 3] == 0.0).all()
    d_x = torch.empty((B, n), dtype=torch.float64, device=dev); d_st = torch.empty(B, dtype=torch.int32, device=dev)
    eng.solve_device(d_blobs, d_x, d_st)
    torch.cuda.synchronize()
    xp, stp, _ = port.port_solve(d, ref)
    ok = (stp == 0)
    assert np.array_equal(d_st.cpu().numpy() == 0, ok)
    assert np.abs(d_x.cpu().numpy() - xp)[ok].max() <= 1e-9
    # argument checks: a dense level cannot receive the postural error
    with pytest.raises(Exception):
        eng.produce_posture_device(d_blobs, t(q), q_ref=t(q_ref), postural_level=1)


def test_batched_math_of_tasks_frontend(ob, port):
    """the pyopensot-like front-end over CUDA tensors (opensot_b200/pysot.py). The
    C29 stack written with the reference's operators -- (com / (0.1*l_wrist + r_wrist) / postural)
    << joint_limits << vel_limits, the shape of coman_ik.cpp:442 -- piled on the device and solved
    in one launch equals the oracle on the same numbers; so do a sub-task (%) and a task turned
    into an equality (<< task)."""
    import torch
    from opensot_b200 import pysot as S
    d = port.shape("C29")
    L = ob.blob_layout(d)
    B, n = 256, d.n
    blobs = port.synth(d, 41, 0, B, 0.15)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A0 = blobs[:, L.A[0]:L.A[0] + 3 * n].reshape(B, 3, n); b0 = blobs[:, L.b[0]:L.b[0] + 3]
    A1 = blobs[:, L.A[1]:L.A[1] + 12 * n].reshape(B, 12, n); b1 = blobs[:, L.b[1]:L.b[1] + 12]
    b2 = blobs[:, L.b[2]:L.b[2] + n]
    com = S.GenericTask("com", t(A0), t(b0))
    l_wrist = S.GenericTask("l_wrist", t(A1[:, :6]), t(b1[:, :6]))
    r_wrist = S.GenericTask("r_wrist", t(A1[:, 6:]), t(b1[:, 6:]))
    postural = S.IdentityTask("postural", t(b2))
    jl = S.GenericConstraint("joint_limits", t(blobs[:, L.u[0]:L.u[0] + n]), t(blobs[:, L.l[0]:L.l[0] + n]))
    vl = S.GenericConstraint("vel_limits", t(blobs[0, L.u[1]:L.u[1] + n]), t(blobs[0, L.l[1]:L.l[1] + n]))
    stack = (com / (0.1 * l_wrist + r_wrist) / postural) << jl << vl
    solver = S.iHQP(stack)
    assert solver.getNumberOfTasks() == 3 and solver.getBackEndName().startswith("B200:rr_solve_kernel")
    ok, x = solver.solve()
    ref = blobs.copy()
    ref[:, L.w[1]:L.w[1] + 12] = np.r_[0.1 * np.ones(6), np.ones(6)]
    ref[:, L.u[1]:L.u[1] + n] = blobs[0, L.u[1]:L.u[1] + n]; ref[:, L.l[1]:L.l[1] + n] = blobs[0, L.l[1]:L.l[1] + n]
    xp, stp, _ = port.port_solve(d, ref)
    assert ok and (stp == 0).all()
    assert np.abs(x.cpu().numpy() - xp).max() <= 1e-9
    # setActiveStack: level 1 off (iHQP.cpp:296-310)
    solver.setActiveStack(1, False)
    ok, x = solver.solve()
    xo, sto, _ = port.port_solve(d, ref, active=[True, False, True])
    assert ok and np.abs(x.cpu().numpy() - xo).max() <= 1e-9
    solver.activateAllStacks()
    # sub-task rows and an equality: ((com % [2, 0]) / postural) << jl << (r_wrist % [0, 1, 2])
    st2 = ((com % [2, 0]) / postural) << jl << (r_wrist % [0, 1, 2])
    s2 = S.iHQP(st2)
    ok2, x2 = s2.solve()
    d2 = _abi.make_desc(n, [dict(m=2), dict(kind=_abi.TASK_IDENTITY)], c_eq=3, n_bound_sets=1)
    L2 = ob.blob_layout(d2)
    r2 = np.zeros((B, L2.stride))
    r2[:, L2.A[0]:L2.A[0] + 2 * n] = A0[:, [2, 0]].reshape(B, -1); r2[:, L2.b[0]:L2.b[0] + 2] = b0[:, [2, 0]]
    r2[:, L2.b[1]:L2.b[1] + n] = b2
    r2[:, L2.Ceq:L2.Ceq + 3 * n] = A1[:, 6:9].reshape(B, -1); r2[:, L2.deq:L2.deq + 3] = b1[:, 6:9]
    r2[:, L2.l[0]:L2.l[0] + n] = blobs[:, L.l[0]:L.l[0] + n]; r2[:, L2.u[0]:L2.u[0] + n] = blobs[:, L.u[0]:L.u[0] + n]
    xp2, stp2, _ = port.port_solve(d2, r2)
    good = stp2 == 0
    assert np.array_equal(s2.getStatus().cpu().numpy() == 0, good)
    assert np.abs(x2.cpu().numpy() - xp2)[good].max() <= 1e-8


def test_math_of_tasks_frontend_per_instance_weights(ob, port):
    """weights with a batch axis (pysot.py: torch tensors [B], [B, m], [B, m, m] are per instance): every instance of the
    C29 stack with its own soft priority of the left wrist (what a policy would output), its own diagonal on the right
    wrist and on the postural level; the weights are copied into the blobs on the device, solved in one launch, and every
    instance equals the oracle on a blob carrying ITS weights. CPU twin: tests/test_pysot_plan.py (bit-exact piling)."""
    import torch
    from opensot_b200 import pysot as S
    d = port.shape("C29")
    L = ob.blob_layout(d)
    B, n = 256, d.n
    blobs = port.synth(d, 43, 0, B, 0.15)
    rng = np.random.default_rng(17)
    w_l = 0.05 + rng.random(B); w_r = 0.5 + rng.random((B, 6)); w_p = 0.5 + rng.random((B, n))
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A0 = blobs[:, L.A[0]:L.A[0] + 3 * n].reshape(B, 3, n); b0 = blobs[:, L.b[0]:L.b[0] + 3]
    A1 = blobs[:, L.A[1]:L.A[1] + 12 * n].reshape(B, 12, n); b1 = blobs[:, L.b[1]:L.b[1] + 12]
    b2 = blobs[:, L.b[2]:L.b[2] + n]
    com = S.GenericTask("com", t(A0), t(b0))
    l_wrist = S.GenericTask("l_wrist", t(A1[:, :6]), t(b1[:, :6]))
    r_wrist = S.GenericTask("r_wrist", t(A1[:, 6:]), t(b1[:, 6:]))
    postural = S.IdentityTask("postural", t(b2))
    jl = S.GenericConstraint("joint_limits", t(blobs[:, L.u[0]:L.u[0] + n]), t(blobs[:, L.l[0]:L.l[0] + n]))
    vl = S.GenericConstraint("vel_limits", t(blobs[:, L.u[1]:L.u[1] + n]), t(blobs[:, L.l[1]:L.l[1] + n]))
    stack = (com / (t(w_l) * l_wrist + t(w_r) * r_wrist) / (t(w_p) * postural)) << jl << vl
    solver = S.iHQP(stack)
    ok, x = solver.solve()
    d2 = _abi.make_desc(n, [dict(m=3), dict(m=12, wkind=_abi.W_DIAG), dict(kind=_abi.TASK_IDENTITY, wkind=_abi.W_DIAG)], n_bound_sets=2)
    assert bytes(solver.desc) == bytes(d2)
    L2 = ob.blob_layout(d2)
    ref = np.zeros((B, L2.stride))
    ref[:, L2.A[0]:L2.A[0] + 3 * n] = A0.reshape(B, -1); ref[:, L2.b[0]:L2.b[0] + 3] = b0
    ref[:, L2.A[1]:L2.A[1] + 12 * n] = A1.reshape(B, -1); ref[:, L2.b[1]:L2.b[1] + 12] = b1
    ref[:, L2.w[1]:L2.w[1] + 12] = np.c_[w_l[:, None] * np.ones((B, 6)), w_r]
    ref[:, L2.b[2]:L2.b[2] + n] = b2; ref[:, L2.w[2]:L2.w[2] + n] = w_p
    for s_ in range(2):
        ref[:, L2.l[s_]:L2.l[s_] + n] = blobs[:, L.l[s_]:L.l[s_] + n]; ref[:, L2.u[s_]:L2.u[s_] + n] = blobs[:, L.u[s_]:L.u[s_] + n]
    assert np.array_equal(solver.d_blobs.cpu().numpy(), ref)          # piled on the device, bit for bit
    xp, stp, _ = port.port_solve(d2, ref)
    assert ok and (stp == 0).all()
    assert np.abs(x.cpu().numpy() - xp).max() <= 1e-9
    # other weights: other answers (the last level is the one whose residual is not zero on this stack, so its weights
    # decide where the posture error goes; levels that are met exactly do not feel theirs)
    w_p2 = t(0.5 + rng.random((B, n)))
    s2 = S.iHQP((com / (t(w_l) * l_wrist + t(w_r) * r_wrist) / (w_p2 * postural)) << jl << vl)
    ok2, x2 = s2.solve()
    ref[:, L2.w[2]:L2.w[2] + n] = w_p2.cpu().numpy()
    xp2, stp2, _ = port.port_solve(d2, ref)
    assert ok2 and np.abs(x2.cpu().numpy() - xp2).max() <= 1e-9 and np.abs(xp2 - xp).max() > 1e-4


def test_math_of_tasks_frontend_level_features(ob, port):
    """the rest of the front-end: a dense weight (W * task, AutoStack.h:127), a
    task-local constraint and a task-local bound (task << c: that level only, AutoStack.cpp:291-303),
    a task-local equality (task << task), an aggregate with an identity part (com + postural), a
    weighted identity level and a user regularisation task with a diagonal weight
    (AutoStack.h:78-90, iHQP.cpp:274-278). The stack is written with the operators, piled on the
    device, and compared with the oracle on a blob filled by hand with the same numbers."""
    import torch
    from opensot_b200 import pysot as S
    rng = np.random.default_rng(5)
    B, n = 128, 14
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A0, b0 = rng.standard_normal((B, 3, n)), 0.02 * rng.standard_normal((B, 3))
    A1, b1 = rng.standard_normal((B, 4, n)), 0.05 * rng.standard_normal((B, 4))
    bp = 0.02 * rng.standard_normal((B, n))
    Ar, br = rng.standard_normal((B, 2, n)), 0.01 * rng.standard_normal((B, 2))
    Q = rng.standard_normal((4, 4)); W1 = Q @ Q.T + 4 * np.eye(4)
    Cl0 = rng.standard_normal((B, 2, n)); lo0 = -0.01 - 0.01 * rng.random((B, 2)); hi0 = 0.01 + 0.01 * rng.random((B, 2))
    Ce, be = rng.standard_normal((B, 1, n)), 0.005 * rng.standard_normal((B, 1))
    lb1, ub1 = -0.012 * np.ones(n), 0.015 * np.ones(n)       # tighter than the global box: active at level 1 only
    lo, hi = -0.02 * np.ones((B, n)), 0.02 * np.ones((B, n))
    wp = np.linspace(0.5, 2.0, n)
    com = S.GenericTask("com", t(A0), t(b0))
    arm = S.GenericTask("arm", t(A1), t(b1))
    postural = S.IdentityTask("postural", t(bp))
    reg = S.GenericTask("reg", t(Ar), t(br)); reg.setWeight([0.3, 0.7])
    hull = S.GenericConstraint("hull", t(hi0), t(lo0), t(Cl0))
    slow = S.GenericConstraint("slow", t(ub1), t(lb1))
    pin = S.GenericTask("pin", t(Ce), t(be))
    lvl0 = (com << hull) << pin
    lvl1 = (W1 * arm) << slow
    stack = (lvl0 / lvl1 / (wp * postural)) << S.GenericConstraint("vel", t(hi), t(lo))
    stack.setRegularisationTask(reg)
    solver = S.iHQP(stack)
    assert solver.getBackEndName() == "B200:sot_solve_kernel"
    ok, x = solver.solve()
    d = _abi.make_desc(n, [dict(m=3, c_loc=3), dict(m=4, wkind=_abi.W_DENSE, c_loc=n),
                           dict(kind=_abi.TASK_IDENTITY, wkind=_abi.W_DIAG)],
                       n_bound_sets=1, m_reg=2, reg_weight_kind=_abi.W_DIAG)
    L = ob.blob_layout(d)
    r = np.zeros((B, L.stride))
    r[:, L.A[0]:L.A[0] + 3 * n] = A0.reshape(B, -1); r[:, L.b[0]:L.b[0] + 3] = b0
    r[:, L.Cl[0]:L.Cl[0] + 3 * n] = np.concatenate([Cl0, Ce], axis=1).reshape(B, -1)
    r[:, L.lCl[0]:L.lCl[0] + 3] = np.c_[lo0, be]; r[:, L.uCl[0]:L.uCl[0] + 3] = np.c_[hi0, be]
    r[:, L.A[1]:L.A[1] + 4 * n] = A1.reshape(B, -1); r[:, L.b[1]:L.b[1] + 4] = b1
    r[:, L.w[1]:L.w[1] + 16] = W1.reshape(-1)
    r[:, L.Cl[1]:L.Cl[1] + n * n] = np.eye(n).reshape(-1)
    r[:, L.lCl[1]:L.lCl[1] + n] = lb1; r[:, L.uCl[1]:L.uCl[1] + n] = ub1
    r[:, L.b[2]:L.b[2] + n] = bp; r[:, L.w[2]:L.w[2] + n] = wp
    r[:, L.Ar:L.Ar + 2 * n] = Ar.reshape(B, -1); r[:, L.br:L.br + 2] = br; r[:, L.wr:L.wr + 2] = [0.3, 0.7]
    r[:, L.l[0]:L.l[0] + n] = lo; r[:, L.u[0]:L.u[0] + n] = hi
    assert bytes(solver.desc) == bytes(d)
    assert np.array_equal(solver.d_blobs.cpu().numpy(), r)       # piling is exact
    xp, stp, _ = port.port_solve(d, r)
    good = stp == 0
    assert good.sum() >= B // 2
    assert np.array_equal(solver.getStatus().cpu().numpy() == 0, good) and ok == bool(good.all())
    assert np.abs(x.cpu().numpy() - xp)[good].max() <= 1e-8
    # an aggregate with an identity part: (com + postural) as one level (tasks/Aggregated.cpp:113-132)
    com2 = S.GenericTask("com", t(A0), t(b0))      # `com` carries its local constraints into an aggregate
    s2 = S.iHQP(((com2 + postural) / arm) << S.GenericConstraint("vel", t(hi), t(lo)))
    ok2, x2 = s2.solve()
    d2 = _abi.make_desc(n, [dict(m=3 + n), dict(m=4)], n_bound_sets=1)
    L2 = ob.blob_layout(d2)
    r2 = np.zeros((B, L2.stride))
    r2[:, L2.A[0]:L2.A[0] + (3 + n) * n] = np.concatenate([A0, np.broadcast_to(np.eye(n), (B, n, n))], axis=1).reshape(B, -1)
    r2[:, L2.b[0]:L2.b[0] + 3 + n] = np.c_[b0, bp]
    r2[:, L2.A[1]:L2.A[1] + 4 * n] = A1.reshape(B, -1); r2[:, L2.b[1]:L2.b[1] + 4] = b1
    r2[:, L2.l[0]:L2.l[0] + n] = lo; r2[:, L2.u[0]:L2.u[0] + n] = hi
    xp2, stp2, _ = port.port_solve(d2, r2)
    assert ok2 and (stp2 == 0).all()
    assert np.abs(x2.cpu().numpy() - xp2).max() <= 1e-8


def test_every_register_row_instance(ob, port):
    """one stack per register-row kernel instance -- 30 columns x 1 row per lane (C29), 30 x 2
    (24 DoF with 14 one-sided rows: 38 rows of y), 38 x 2 (C38c) -- and one per reason to stay on
    the generic kernel (narrow, wide, dense weight): the engine picks the documented kernel and
    the answers match the live reference / the restatement."""
    import os
    d24 = _abi.make_desc(24, [dict(m=6), dict(m=8, wkind=_abi.W_DIAG), dict(kind=_abi.TASK_IDENTITY)],
                         c_in=14, n_bound_sets=2)
    d48 = _abi.make_desc(48, [dict(m=6), dict(m=12, wkind=_abi.W_DIAG), dict(kind=_abi.TASK_IDENTITY)],
                         c_in=10, n_bound_sets=2)
    # (descriptor, OSB_KERNEL override, kernel the engine must pick): up to 32 rows of y -> one warp per stack;
    # 33..96 rows or 39..64 variables -> one CTA of 2-3 warps per stack (the two-rows-per-lane register-row
    # instances stay reachable with OSB_KERNEL=rr); narrow stacks and regularisation tasks -> generic kernel; a dense weight
    # on a dense task is whitened in place and stays on the fast kernels
    cases = [(port.shape("C29"), None, "rr_solve_kernel<f64,30x1>"),
             (d24, None, "mw_solve_kernel<f64,38x2 warps>"), (d24, "rr", "rr_solve_kernel<f64,30x2>"),
             (port.shape("C38c"), None, "mw_solve_kernel<f64,38x2 warps>"), (port.shape("C38c"), "rr", "rr_solve_kernel<f64,38x2>"),
             (port.shape("C35f"), None, "mw_solve_kernel<f64,38x2 warps>"),
             (d48, None, "mw_solve_kernel<f64,64x2 warps>"), (port.shape("C64c"), None, "mw_solve_kernel<f64,64x3 warps>"),
             (port.shape("C64c"), "generic", "sot_solve_kernel"),
             (port.shape("C12"), None, "sot_solve_kernel"), (port.shape("C29r"), None, "sot_solve_kernel"),
             (port.shape("C29w"), None, "rr_solve_kernel<f64,30x1>"), (port.shape("C29w"), "generic", "sot_solve_kernel")]
    for d, force, kernel in cases:
        B = 96
        if force:
            os.environ["OSB_KERNEL"] = force
        try:
            eng = ob.BatchedStackSolver(d, B)
        finally:
            os.environ.pop("OSB_KERNEL", None)
        assert eng.kernel_name() == kernel, (eng.kernel_name(), kernel)
        blobs = port.synth(d, 77, 0, B, 0.2)
        x, st = eng.solve_host(blobs)
        if port.have_ref():
            ok, xr, _ = port.ref_solve(d, blobs)
        else:
            xr, stp, _ = port.port_solve(d, blobs); ok = stp == 0
        xp, stp, _ = port.port_solve(d, blobs)
        assert not ((st != 0) & ok).any(), kernel
        stable = (st == 0) & ok & (stp == 0) & (np.abs(xr - xp).max(axis=1) <= common.STABLE_TOL)
        assert stable.sum() >= 0.9 * B
        assert np.abs(x - xr)[stable].max() <= common.PARITY_TOL, kernel
        eng.close()


def test_fp32_mpc_config(ob, port):
    """BASELINE configs[3]: MPC-style 29-DoF stack x horizon 20 x 512 rollouts = 10240 instances,
    fp32. The reference has no fp32 path, so the arbiter is the fp64 oracle on the same inputs.
    STATED TOLERANCE (fp32 mode): joint-velocity infinity-norm <= 1e-5 on this workload
    (velocities are O(1e-2); fp32 unit round-off 6e-8 times the conditioning of the cascade)."""
    x, st, xp, stp = _f32_errors(ob, port, "C29", 0.10, 10240, 4321)
    assert (st == 0).all() and (stp == 0).all()
    err = np.abs(x - xp).max(axis=1)
    print("fp32 C29 x 10240: max %.2e p99 %.2e median %.2e" % (err.max(), np.percentile(err, 99), np.median(err)))
    assert err.max() <= 1e-5
    assert np.median(err) <= 1e-7


@pytest.mark.parametrize("shape,tight,B", [("C29", 0.35, 1024), ("C35f", 0.1, 512), ("C38c", 0.1, 512), ("C12", 0.3, 512), ("C64c", 0.1, 64)])
def test_fp32_other_shapes(ob, port, shape, tight, B):
    """fp32 on harder / larger shapes. Many saturated bounds make the last level ill-conditioned
    (the same instances on which qpOASES itself moves by 1e-4), so the
    stated bound is: all instances within 1e-4 (0.5 % of the velocity bound), 99 % within 1e-5 (measured on B200: worst
    shape max 9.1e-6, p99 2.3e-6; round 1 stated 2e-3 / 1e-4)."""
    x, st, xp, stp = _f32_errors(ob, port, shape, tight, B, 99)
    good = (st == 0) & (stp == 0)
    assert good.mean() >= 0.98
    err = np.abs(x - xp).max(axis=1)[good]
    print("fp32 %s tight=%.2f: max %.2e p99 %.2e median %.2e" % (shape, tight, err.max(), np.percentile(err, 99), np.median(err)))
    assert np.percentile(err, 99) <= 1e-5 and err.max() <= 1e-4 and np.median(err) <= 1e-6


def test_reference_stack_kats(ob, port):
    """the stack-path known-answer tests of the reference (tests/test_stack_kats.py: TestGenericTask.cpp:175-203,
    TestBasicAlgebra.cpp:46-108, examples/cpp/simple_qp.cpp) through the C ABI: the reference's own
    assertion holds for the GPU answer; on the rank-deficient 12 x 30 case the GPU reproduces the
    exact minimiser to round-off (qpOASES, which factors H = A'A + eps I, is ~1e-5 away)."""
    import test_stack_kats as K
    for name, case in K.CASES:
        d, fill, check = case[:3]
        extra = case[3] if len(case) > 3 else {}
        L, blob = K.build(ob.blob_layout, d, fill)
        eng = ob.BatchedStackSolver(d, 4)
        x, st = eng.solve_host(np.repeat(blob, 4, axis=0))
        assert (st == 0).all() and np.array_equal(x[0], x[3]), name
        assert check(x[0], blob[0], L), name
        xp, stp, _ = port.port_solve(d, blob)
        assert stp[0] == 0 and np.abs(x[0] - xp[0]).max() <= extra.get("xtol", common.PARITY_TOL), name
        if "exact" in extra:
            assert np.abs(x[0] - extra["exact"]).max() <= 1e-9, name
        eng.close()


@pytest.mark.parametrize("name,count", [("C29", 24), ("C38c", 12), ("C35f", 12), ("C29w", 8), ("C29r", 8)])
def test_logging_from_cuda_outputs(ob, port, tmp_path, name, count):
    """SURVEY 8f rank 4 on the CUDA path: BatchedStackSolver.log writes what iHQP::log / BackEnd::log dump
    (src/solvers/iHQP.cpp:403-415, BackEnd.cpp:111-130) from the blobs and the per-level solutions the KERNELS
    returned (store_levels); every logged level QP, handed to the oracle's single-QP solver with the back-end's eps,
    reproduces the logged solution_i; the last level's solution is x."""
    d = port.shape(name)
    blobs = port.synth(d, 909, 0, count, 0.2)
    eng = ob.BatchedStackSolver(d, count, store_levels=True)
    x, st, xl = eng.solve_host(blobs)
    assert (st == 0).all()
    log = eng.log(str(tmp_path / "cuda_log.npz"), blobs, xl, task_ids=["t%d" % k for k in range(d.n_levels)])
    z = np.load(tmp_path / "cuda_log.npz")
    assert np.array_equal(z["solution_%d" % (d.n_levels - 1)], x) and "t0_A" in z
    checked, worst = 0, 0.0
    for i in range(count):
        for k in range(d.n_levels):
            A = log.get("A_%d" % k)
            xk, stk, _ = port.port_qp(log["H_%d" % k][i], log["g_%d" % k][i], None if A is None else A[i],
                                      None if A is None else log["lA_%d" % k][i], None if A is None else log["uA_%d" % k][i],
                                      log["l_%d" % k][i] if d.n_bound_sets else None,
                                      log["u_%d" % k][i] if d.n_bound_sets else None, eps=d.eps)
            if stk != 0:       # dependent equality rows in the oracle's Goldfarb-Idnani QP: no verdict
                continue
            # an earlier level fixes A_k x only (the rest of its x is held by eps): compare in task space there
            if k < d.n_levels - 1 and d.task_kind[k] == _abi.TASK_DENSE:
                L = ob.blob_layout(d)
                Ak = blobs[i, L.A[k]:L.A[k] + d.m[k] * d.n].reshape(d.m[k], d.n)
                e = np.abs(Ak @ (xk - log["solution_%d" % k][i])).max()
            else:
                e = np.abs(xk - log["solution_%d" % k][i]).max()
            worst = max(worst, e)
            checked += 1
    print("%s: %d logged level QPs of the CUDA path re-solved by the oracle, worst %.2e" % (name, checked, worst))
    assert checked >= count and worst <= common.PARITY_TOL


@pytest.mark.parametrize("shape,B", [("C29", 4096), ("C38c", 2048)])
def test_fp32_against_the_reference_at_its_default_eps(ob, port, shape, B):
    """What the fp32 engine costs in accuracy against the REFERENCE AS SHIPPED: the fp32 engine raises eps to 1e-6
    (OSB_F32_MIN_EPS: the default 4.4e-11 is below single-precision resolution), the live reference (fp64 qpOASES)
    keeps its default eps factor 2e2. Reported, with a stated bound: p99 <= 1e-5 and max <= 1e-4 on the
    reference-stable set (fp64 engines: 1e-6; measured on B200: max 1.2e-6). The distance has two parts, printed apart: the eps change itself
    (fp64 engine at eps 1e-6 vs the reference) and single precision (fp32 vs fp64 engine at the same eps)."""
    if not port.have_ref():
        pytest.fail("oracle/_ref missing on the GPU box")
    d = port.shape(shape)                       # default eps factor 2e2
    blobs = port.synth(d, 2026, 0, B, 0.10)
    okh, okc, xh, xc = port.ref_solve_both(d, blobs)
    L = ob.blob_layout(d)
    stable, _, _ = common.live_stable_mask(d, L, blobs, okh, okc, xh, xc)
    e32 = ob.BatchedStackSolver(d, B, precision="f32")
    x32, st32 = e32.solve_host(e32.pack(blobs))
    d6 = port.shape(shape, _abi.F32_EPS_FACTOR)
    e64 = ob.BatchedStackSolver(d6, B)
    x64, st64 = e64.solve_host(blobs)
    good = stable & (st32 == 0) & (st64 == 0)
    err = np.abs(x32 - xh).max(axis=1)[good]
    eps_part = np.abs(x64 - xh).max(axis=1)[good]
    f32_part = np.abs(x32 - x64).max(axis=1)[good]
    print("fp32 %s x %d vs live reference at default eps: max %.2e p99 %.2e median %.2e | eps 4.4e-11 -> 1e-6 alone: max %.2e | "
          "single precision alone: max %.2e p99 %.2e" % (shape, B, err.max(), np.percentile(err, 99), np.median(err), eps_part.max(),
                                                       f32_part.max(), np.percentile(f32_part, 99)))
    assert good.mean() >= 0.97
    assert np.percentile(err, 99) <= 1e-5 and err.max() <= 1e-4


def test_engines_of_different_shapes_share_a_kernel(ob, port):
    """the opt-in shared-memory size belongs to the kernel FUNCTION: a small engine created after a large one on the
    same kernel must not shrink it (found by the nHQP front-end, three engines on sot_solve_kernel)"""
    big, small = port.shape("C29r"), port.shape("C12")      # (user regularisation task / narrow stack: both on the generic kernel)
    eb = ob.BatchedStackSolver(big, 32)
    es = ob.BatchedStackSolver(small, 32)
    assert eb.kernel_name() == es.kernel_name() == "sot_solve_kernel"
    bb, bs = port.synth(big, 4, 0, 32, 0.2), port.synth(small, 4, 0, 32, 0.2)
    for eng, d, blobs in ((es, small, bs), (eb, big, bb), (es, small, bs)):
        x, st = eng.solve_host(blobs)
        xp, stp, _ = port.port_solve(d, blobs)
        assert (st == 0).all() and np.abs(x - xp).max() <= common.PARITY_TOL


@pytest.mark.parametrize("name,tight,B", [("C29c", 0.5, 1024), ("C29c", 0.8, 512), ("C29", 0.8, 1024), ("C35f", 0.8, 512),
                                          ("C38c", 0.6, 512), ("C35h", 0.6, 512)])
def test_no_false_infeasible_and_kernels_agree_when_saturated(ob, port, name, tight, B):
    """(i) linear cost terms on rank-deficient levels / heavily saturated stacks used to end in a
    false OSB_INFEASIBLE at fully degenerate vertices: now no instance fails that the live reference solves; (ii) the
    kernel the engine picks and the generic kernel (another data layout, same mathematics) agree on every status word
    and on the answers, without any tolerance constant added."""
    import os
    d = port.shape(name)
    blobs = port.synth(d, 7001, 0, B, tight)
    ea = ob.BatchedStackSolver(d, B)
    os.environ["OSB_KERNEL"] = "generic"
    try:
        eg = ob.BatchedStackSolver(d, B)
    finally:
        os.environ.pop("OSB_KERNEL", None)
    assert ea.kernel_name() != eg.kernel_name()
    xa, sa = ea.solve_host(blobs)
    xg, sg = eg.solve_host(blobs)
    sub = slice(0, 192)
    ok, xr, _ = port.ref_solve(d, blobs[sub]) if port.have_ref() else (np.ones(192, bool), xa[sub], None)
    L = ob.blob_layout(d)
    both = (sa == 0) & (sg == 0)
    print("%s tight %.1f x %d: %s fails %d, generic fails %d, status mismatches %d, max |x_a - x_g| %.2e; fails where the reference "
          "solves: %d of %d checked; max violation %.1e" % (name, tight, B, ea.kernel_name(), (sa != 0).sum(), (sg != 0).sum(),
          (sa != sg).sum(), np.abs(xa - xg)[both].max(), ((sa[sub] != 0) & ok).sum(), ok.sum(),
          common.max_violation(d, L, blobs[sa == 0], xa[sa == 0]).max()))
    assert np.array_equal(sa != 0, sg != 0)
    # C29c: a linear cost term on a rank-deficient level is an LP in that level's null space, held by eps alone (the
    # unconstrained minimiser sits |c| / eps away): two correct elimination orders differ by cond * u there, the CPU
    # arbiters themselves by 1e-7 (test_every_level_path_against_reference); elsewhere the parity bound holds
    assert np.abs(xa - xg)[both].max() <= (1e-5 if name == "C29c" else common.PARITY_TOL)
    assert not ((sa[sub] != 0) & ok).any()
    assert common.max_violation(d, L, blobs[sa == 0], xa[sa == 0]).max() <= 1e-9


def test_two_host_calls_in_flight(ob, port):
    """osb_solve_host_submit / osb_solve_wait (throughput mode: two sets of device buffers, the blobs of call k+1 cross
    PCIe while call k is solved): the same bits as the blocking call, for both precisions and for ragged batch sizes;
    a third submit, a wait without a submit and a blocking call while calls are in flight are refused loudly."""
    import torch
    d = port.shape("C29")
    B = 700
    sets = [port.synth(d, 4242 + s, 0, B, 0.15) for s in range(5)]
    for prec in ("f64", "f32"):
        eng = ob.BatchedStackSolver(d, B, precision=prec)
        packed = [torch.from_numpy(eng.pack(b)).pin_memory() for b in sets]
        want = [eng.solve_host(p.numpy()) for p in packed]
        xs = [torch.empty((B, d.n), dtype=packed[0].dtype).pin_memory() for _ in range(2)]
        sts = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
        with pytest.raises(RuntimeError):
            eng.wait()
        got = []
        for i, p in enumerate(packed):
            cnt = B - 37 * i                         # ragged: every call its own batch size
            if eng.in_flight() == 2:
                eng.wait()
                got.append((xs[i & 1].numpy()[:B - 37 * (i - 2)].copy(), sts[i & 1].numpy()[:B - 37 * (i - 2)].copy()))
            eng.submit_host(p.numpy()[:cnt], xs[i & 1].numpy(), sts[i & 1].numpy())
        assert eng.in_flight() == 2
        with pytest.raises(RuntimeError):
            eng.submit_host(packed[0].numpy(), xs[0].numpy(), sts[0].numpy())
        with pytest.raises(RuntimeError):
            eng.solve_host(packed[0].numpy())
        for i in (len(packed) - 2, len(packed) - 1):
            eng.wait()
            got.append((xs[i & 1].numpy()[:B - 37 * i].copy(), sts[i & 1].numpy()[:B - 37 * i].copy()))
        assert eng.in_flight() == 0
        for i, (x, st) in enumerate(got):
            cnt = B - 37 * i
            assert np.array_equal(st, want[i][1][:cnt]) and np.array_equal(x, want[i][0][:cnt]), (prec, i)
        x_again, _ = eng.solve_host(packed[1].numpy())       # the blocking call works again afterwards
        assert np.array_equal(x_again, want[1][0])
        eng.close()


@pytest.mark.parametrize("shape,prec", [("C29", "f64"), ("C29", "f32"), ("C35f", "f64"), ("C38c", "f64"), ("C35h", "f64")])
def test_wire_format(ob, port, shape, prec):
    """osb_wire_configure / osb_solve_host_wire(_submit): only the fields that vary cross PCIe, the static ones come from a
    template on the device. Same bits as the complete-blob call (the kernel reads identical blobs), blocking and with two
    calls in flight; a complete-blob call or a state-input call in between does not leave its data in the static fields."""
    import torch
    from opensot_b200 import synth
    d = port.shape(shape) if prec == "f64" else port.shape(shape, _abi.F32_EPS_FACTOR)
    B = 300
    eng = ob.BatchedStackSolver(d, B, precision=prec)
    L = ob.blob_layout(d)
    segs = synth.static_segments(d, L)
    sets = [port.synth(d, 31 + s, 1000 * s, B, 0.2) for s in range(3)]
    packed = [eng.pack(b) for b in sets]
    want = [eng.solve_host(p) for p in packed]
    ws = eng.wire_configure(segs, sets[0][0])
    assert ws == eng.wire_stride() and ws < eng.stride and ws >= L.stride - sum(c for _, c in segs)
    wires = [torch.from_numpy(eng.wire_pack(p)).pin_memory() for p in packed]
    x, st = eng.solve_host_wire(wires[0].numpy())
    assert np.array_equal(x, want[0][0]) and np.array_equal(st, want[0][1])
    # a complete-blob call with OTHER static values in between: the next wire call must not see them
    other = packed[1].copy()
    for off, cnt in segs:
        other[:, off:off + cnt] *= 0.5
    eng.solve_host(other)
    x, st = eng.solve_host_wire(wires[1].numpy())
    assert np.array_equal(x, want[1][0]) and np.array_equal(st, want[1][1])
    # two in flight (the second slot gets its own template)
    xs = [torch.empty((B, d.n), dtype=wires[0].dtype).pin_memory() for _ in range(2)]
    sts = [torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(2)]
    eng.submit_host_wire(wires[2].numpy(), xs[0].numpy(), sts[0].numpy())
    eng.submit_host_wire(wires[0].numpy()[:B - 11], xs[1].numpy(), sts[1].numpy())
    eng.wait(); eng.wait()
    assert np.array_equal(xs[0].numpy(), want[2][0]) and np.array_equal(sts[0].numpy(), want[2][1])
    assert np.array_equal(xs[1].numpy()[:B - 11], want[0][0][:B - 11])
    # nothing static: refused; overlapping segments: refused
    with pytest.raises(RuntimeError):
        eng.wire_configure([(0, L.stride)], sets[0][0])
    with pytest.raises(RuntimeError):
        eng.wire_configure([(0, 10), (5, 10)], sets[0][0])
    eng.close()

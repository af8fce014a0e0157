"""CPU suite: host logic of the Math-of-Tasks front-end (opensot_b200/pysot.py) without a GPU.
`iHQP(stack, plan_only=True)` derives the stack descriptor, the blob layout and the piling plan
(the list of gather sources osb_stack_gather_device executes); the plan is run here by a numpy
restatement of the gather kernel (sot_gather_kernel: dst[r*cols+c] = scale * src[row_index[r]*ld+c])
and must produce, bit for bit, the blob written by hand from the same numbers -- for the stacks of
the GPU tests (tests/test_gpu_parity.py: operators + / % * <<, dense and diagonal weights, task-local
constraints / bounds / equalities, aggregates with identity parts, regularisation task). The oracle
then solves the piled blob (no CPU solve path in the product: solve() raises in plan-only mode)."""
import numpy as np
import pytest
import torch

from opensot_b200 import _abi, pysot as S


def run_plan(solver):
    """numpy restatement of osb_stack_gather_device over the solver's own piling plan"""
    srcs = solver._sources()                     # also writes the (constant) weights into the blob buffer
    blobs = solver.d_blobs.numpy().copy()
    for s in srcs:
        src = s["src"]
        rows, cols, ld = s["rows"], s["cols"], s.get("src_ld", s["cols"])
        idx = s.get("row_index")
        ridx = np.arange(rows) if idx is None else idx.numpy()
        base = src.numpy() if src.is_contiguous() else None
        for i in range(blobs.shape[0]):
            flat = (src[i].contiguous().numpy() if base is None else base[i]).reshape(-1)
            block = np.stack([flat[r * ld:r * ld + cols] for r in ridx])
            blobs[i, s["dst_off"]:s["dst_off"] + rows * cols] = (s.get("scale", 1.0) * block).reshape(-1)
    return blobs


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_coman_like_stack_plan(port):
    """(com / (0.1*l_wrist + r_wrist) / postural) << joint_limits << vel_limits, coman_ik.cpp:442"""
    d = port.shape("C29")
    L = port.layout(d)
    B, n = 8, d.n
    blobs = port.synth(d, 41, 0, B, 0.15)
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
    solver = S.iHQP(stack, plan_only=True)
    assert bytes(solver.desc) == bytes(d) and solver.getNumberOfTasks() == 3
    ref = blobs.copy()
    ref[:, L.w[1]:L.w[1] + 12] = np.r_[0.1 * np.ones(6), np.ones(6)]
    ref[:, L.u[1]:L.u[1] + n] = blobs[0, L.u[1]:L.u[1] + n]; ref[:, L.l[1]:L.l[1] + n] = blobs[0, L.l[1]:L.l[1] + n]
    assert np.array_equal(run_plan(solver), ref)
    with pytest.raises(RuntimeError):
        solver.solve()
    # sub-task rows and a task as a global equality
    s2 = S.iHQP(((com % [2, 0]) / postural) << jl << (r_wrist % [0, 1, 2]), plan_only=True)
    d2 = _abi.make_desc(n, [dict(m=2), dict(kind=_abi.TASK_IDENTITY)], c_eq=3, n_bound_sets=1)
    L2 = port.layout(d2)
    r2 = np.zeros((B, L2.stride))
    r2[:, L2.A[0]:L2.A[0] + 2 * n] = A0[:, [2, 0]].reshape(B, -1); r2[:, L2.b[0]:L2.b[0] + 2] = b0[:, [2, 0]]
    r2[:, L2.b[1]:L2.b[1] + n] = b2
    r2[:, L2.Ceq:L2.Ceq + 3 * n] = A1[:, 6:9].reshape(B, -1); r2[:, L2.deq:L2.deq + 3] = b1[:, 6:9]
    r2[:, L2.l[0]:L2.l[0] + n] = blobs[:, L.l[0]:L.l[0] + n]; r2[:, L2.u[0]:L2.u[0] + n] = blobs[:, L.u[0]:L.u[0] + n]
    assert bytes(s2.desc) == bytes(d2) and np.array_equal(run_plan(s2), r2)


def test_level_features_plan(port):
    """dense weight, task-local rows / bound / equality, weighted identity level, regularisation
    task, aggregate with an identity part: descriptor and piled blob as written by hand; the
    local bound (tighter than the global box) binds at its level only."""
    rng = np.random.default_rng(5)
    B, n = 16, 14
    A0, b0 = rng.standard_normal((B, 3, n)), 0.02 * rng.standard_normal((B, 3))
    A1, b1 = rng.standard_normal((B, 4, n)), 0.05 * rng.standard_normal((B, 4))
    bp = 0.02 * rng.standard_normal((B, n))
    Ar, br = rng.standard_normal((B, 2, n)), 0.01 * rng.standard_normal((B, 2))
    Q = rng.standard_normal((4, 4)); W1 = Q @ Q.T + 4 * np.eye(4)
    Cl0 = rng.standard_normal((B, 2, n)); lo0 = -0.01 - 0.01 * rng.random((B, 2)); hi0 = 0.01 + 0.01 * rng.random((B, 2))
    Ce, be = rng.standard_normal((B, 1, n)), 0.005 * rng.standard_normal((B, 1))
    lb1, ub1 = -0.012 * np.ones(n), 0.015 * np.ones(n)
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
    assert lvl0 is com and com.getConstraints() == [hull, pin]          # AutoStack.cpp:291-303: the task is returned
    lvl1 = (W1 * arm) << slow
    assert lvl1.weight_kind() == _abi.W_DENSE and np.allclose(lvl1.getWeight(), W1)
    stack = (lvl0 / lvl1 / (wp * postural)) << S.GenericConstraint("vel", t(hi), t(lo))
    stack.setRegularisationTask(reg)
    solver = S.iHQP(stack, plan_only=True)
    d = _abi.make_desc(n, [dict(m=3, c_loc=3), dict(m=4, wkind=_abi.W_DENSE, c_loc=n),
                           dict(kind=_abi.TASK_IDENTITY, wkind=_abi.W_DIAG)],
                       n_bound_sets=1, m_reg=2, reg_weight_kind=_abi.W_DIAG)
    L = port.layout(d)
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
    piled = run_plan(solver)
    assert np.array_equal(piled, r)
    x, xl, st, _ = port.port_solve(d, piled, levels=True)
    assert (st == 0).sum() >= B // 2
    ok = st == 0
    assert (xl[ok, 1] <= ub1 + 1e-9).all() and (xl[ok, 1] >= lb1 - 1e-9).all()    # the local bound holds at level 1 ...
    assert ((x[ok] > ub1 + 1e-6) | (x[ok] < lb1 - 1e-6)).any()                        # ... and not below it
    # aggregate with an identity part; a child's local constraints travel with it (tasks/Aggregated.cpp:134-200)
    agg = com + postural
    assert agg.getConstraints() == [hull, pin] and agg.getTaskSize() == 3 + n
    com2 = S.GenericTask("com", t(A0), t(b0))
    s2 = S.iHQP(((com2 + postural) / arm) << S.GenericConstraint("vel", t(hi), t(lo)), plan_only=True)
    d2 = _abi.make_desc(n, [dict(m=3 + n), dict(m=4)], n_bound_sets=1)
    L2 = port.layout(d2)
    r2 = np.zeros((B, L2.stride))
    r2[:, L2.A[0]:L2.A[0] + (3 + n) * n] = np.concatenate([A0, np.broadcast_to(np.eye(n), (B, n, n))], axis=1).reshape(B, -1)
    r2[:, L2.b[0]:L2.b[0] + 3 + n] = np.c_[b0, bp]
    r2[:, L2.A[1]:L2.A[1] + 4 * n] = A1.reshape(B, -1); r2[:, L2.b[1]:L2.b[1] + 4] = b1
    r2[:, L2.l[0]:L2.l[0] + n] = lo; r2[:, L2.u[0]:L2.u[0] + n] = hi
    assert bytes(s2.desc) == bytes(d2) and np.array_equal(run_plan(s2), r2)


def test_operator_semantics():
    """weights: w * task scales the weight, not A or b; % selects rows of a diagonal and of a dense weight"""
    B, n = 2, 5
    A, b = torch.randn(B, 4, n, dtype=torch.float64), torch.randn(B, 4, dtype=torch.float64)
    task = S.GenericTask("t", A, b)
    w = 0.5 * task
    assert w.A is task.A and np.array_equal(w.getWeight(), 0.5 * np.ones(4)) and task.weight is None
    W = np.arange(16.0).reshape(4, 4) + 4 * np.eye(4)
    task.setWeight(W)
    sub = task % [3, 1]
    assert np.array_equal(sub.getWeight(), W[np.ix_([3, 1], [3, 1])]) and sub.getTaskSize() == 2
    task.setWeight(np.diag([1.0, 2.0, 3.0, 4.0]))                       # a diagonal matrix stays a diagonal weight
    assert task.weight_kind() == _abi.W_DIAG and np.array_equal((task % [2]).getWeight(), [3.0])
    two = (2.0 * (task + S.GenericTask("u", A[:, :1], b[:, :1])))
    assert np.array_equal(two.getWeight(), [2.0, 4.0, 6.0, 8.0, 2.0])
    st = (task / two) << S.GenericConstraint("box", torch.ones(n, dtype=torch.float64), -torch.ones(n, dtype=torch.float64))
    assert len(st.getStack()) == 2 and len(st.getBoundsList()) == 1
    with pytest.raises(RuntimeError):
        S.iHQP(S.AutoStack([]), plan_only=True)                         # iHQP.cpp:24
    sol = S.iHQP(st, eps_regularisation=2e2, plan_only=True)
    assert sol.desc.eps == 2e2 * _abi.OSB_QPOASES_BASE_EPS               # constructor: a factor (QPOasesBackEnd.cpp:57)
    assert sol.setEpsRegularisation(1e-6) and sol.desc.eps == 1e-6      # setter: absolute (QPOasesBackEnd.cpp:358-366)
    assert not sol.setEpsRegularisation(-1.0) and sol.desc.eps == 1e-6


def test_inactive_task_keeps_b_and_zeroes_A(port):
    """Task::setActive(false): A = 0, b unchanged (Task.h:382-386), also for one part of an aggregate;
    the level then only sees eps |x|^2 inside the earlier levels' optimal set (not the same as a
    level switched off with setActiveStack, which is skipped altogether)."""
    rng = np.random.default_rng(9)
    B, n = 4, 6
    A0, b0 = rng.standard_normal((B, 2, n)), rng.standard_normal((B, 2))
    A1, b1 = rng.standard_normal((B, 3, n)), rng.standard_normal((B, 3))
    t0, t1 = S.GenericTask("t0", t(A0), t(b0)), S.GenericTask("t1", t(A1), t(b1))
    box = S.GenericConstraint("box", t(np.ones(n)), t(-np.ones(n)))
    t1.setActive(False)
    assert not t1.getActive() and not (t0 + t1).getActive() and t0.getActive()
    solver = S.iHQP(((t0 + t1) / S.IdentityTask("post", t(np.zeros((B, n))))) << box, plan_only=True)
    L = solver.layout
    piled = run_plan(solver)
    assert np.array_equal(piled[:, L.A[0]:L.A[0] + 2 * n], A0.reshape(B, -1))
    assert not piled[:, L.A[0] + 2 * n:L.A[0] + 5 * n].any()
    assert np.array_equal(piled[:, L.b[0]:L.b[0] + 5], np.c_[b0, b1])
    t1.setActive(True)
    assert np.array_equal(run_plan(solver)[:, L.A[0] + 2 * n:L.A[0] + 5 * n], A1.reshape(B, -1))
    post = S.IdentityTask("post", t(np.zeros((B, n))))
    post.setActive(False)
    with pytest.raises(NotImplementedError):
        S.iHQP(t0 / post, plan_only=True)._sources()


def test_per_instance_weights_plan(port):
    """Weights with a batch axis (torch tensors: [B] a scalar per instance, [B, m] a diagonal, [B, m, m] a dense matrix;
    numbers and numpy arrays stay per stack): `w * task`, aggregates mixing both kinds, `%` rows of a weighted task.
    The piled blob of instance i must equal, bit for bit, the blob of a batch of one built with instance i's numbers and
    PER-STACK weights (the path the other tests pin), and the oracle must solve it."""
    rng = np.random.default_rng(9)
    B, n = 6, 12
    A0, b0 = rng.standard_normal((B, 4, n)), 0.02 * rng.standard_normal((B, 4))
    A1, b1 = rng.standard_normal((B, 3, n)), 0.02 * rng.standard_normal((B, 3))
    A2, b2 = rng.standard_normal((B, 3, n)), 0.02 * rng.standard_normal((B, 3))
    bp = 0.02 * rng.standard_normal((B, n))
    lo, hi = -0.02 * np.ones((B, n)), 0.02 * np.ones((B, n))
    w_scalar = 0.1 + rng.random(B)                       # one scalar per instance (a policy's soft priority)
    w_diag = 0.5 + rng.random((B, 3))                    # a diagonal per instance
    Q = rng.standard_normal((B, 4, 4)); W_dense = Q @ Q.transpose(0, 2, 1) + 4 * np.eye(4)      # a dense SPD matrix per instance
    w_post = 0.5 + rng.random((B, n))

    def build(sl, per_instance):
        """the stack over the instances `sl`; weights as torch tensors with a batch axis, or (batch of one) per stack"""
        k = lambda a: t(a[sl])
        lift = (lambda a: t(a[sl])) if per_instance else (lambda a: a[sl][0])
        com = S.GenericTask("com", k(A0), k(b0))
        l_w = S.GenericTask("l_wrist", k(A1), k(b1))
        r_w = S.GenericTask("r_wrist", k(A2), k(b2))
        post = S.IdentityTask("postural", k(bp))
        lvl0 = lift(W_dense) * (com % [0, 1, 2, 3])
        lvl1 = lift(w_scalar) * l_w + 0.5 * (lift(w_diag) * r_w)          # per-instance scalar + (per-stack scalar x per-instance diagonal)
        lvl1 = lvl1 if per_instance else lvl1
        lvl2 = lift(w_post) * post
        return (lvl0 / lvl1 / lvl2) << S.GenericConstraint("vel", k(hi), k(lo))

    batched = S.iHQP(build(slice(0, B), True), plan_only=True)
    assert batched.desc.weight_kind[0] == _abi.W_DENSE and batched.desc.weight_kind[1] == _abi.W_DIAG and batched.desc.weight_kind[2] == _abi.W_DIAG
    piled = run_plan(batched)
    L = port.layout(batched.desc)
    for i in range(B):
        one = S.iHQP(build(slice(i, i + 1), False), plan_only=True)
        assert bytes(one.desc) == bytes(batched.desc)
        assert np.array_equal(run_plan(one)[0], piled[i]), i
        # and by hand for the weights
        assert np.array_equal(piled[i, L.w[0]:L.w[0] + 16], W_dense[i].reshape(-1))
        assert np.array_equal(piled[i, L.w[1]:L.w[1] + 6], np.r_[w_scalar[i] * np.ones(3), 0.5 * w_diag[i]])
        assert np.array_equal(piled[i, L.w[2]:L.w[2] + n], w_post[i])
    xp, _, stp, _ = port.port_solve(batched.desc, piled, levels=True)
    assert (stp == 0).all()
    # the weights matter: the same stack with instance 0's weights everywhere gives other answers for the other instances
    same_w = piled.copy()
    for off, cnt in ((L.w[0], 16), (L.w[1], 6), (L.w[2], n)):
        same_w[:, off:off + cnt] = piled[0, off:off + cnt]
    xs, _, _, _ = port.port_solve(batched.desc, same_w, levels=True)
    assert np.abs(xs[1:] - xp[1:]).max() > 1e-5 and np.array_equal(xs[0], xp[0])


def test_row_slots_of_a_constraint_plan(port):
    """A general constraint whose row count differs per instance and changes between solves (ConvexHull-like): declared
    with max_rows, refilled in place with setRows(..., count=[B] tensor). The piled blob carries the real rows and null
    rows -inf <= 0 x <= +inf behind them; for every instance the oracle's answer equals the answer of a stack that has
    only that instance's real rows."""
    rng = np.random.default_rng(21)
    B, n, K = 6, 10, 5
    A0, b0 = rng.standard_normal((B, 3, n)), 0.02 * rng.standard_normal((B, 3))
    bp = 0.02 * rng.standard_normal((B, n))
    lo, hi = -0.02 * np.ones((B, n)), 0.02 * np.ones((B, n))
    task = S.GenericTask("t", t(A0), t(b0))
    post = S.IdentityTask("postural", t(bp))
    C = rng.standard_normal((B, 3, n)); ub = 0.0005 + 0.004 * rng.random((B, 3)); lb = -1e20 * np.ones((B, 3))
    hull = S.GenericConstraint("hull", t(ub), t(lb), t(C), max_rows=K)
    solver = S.iHQP((task / post) << S.GenericConstraint("vel", t(hi), t(lo)) << hull, plan_only=True)
    assert solver.desc.c_in == K
    L = port.layout(solver.desc)

    def check(Cn, ubn, lbn, counts):
        piled = run_plan(solver)
        x, _, st, _ = port.port_solve(solver.desc, piled, levels=True)
        assert (st == 0).all()
        binding = 0
        for i in range(B):
            c = int(counts[i])
            rows = piled[i, L.Cin:L.Cin + K * n].reshape(K, n)
            assert np.array_equal(rows[:c], Cn[i, :c]) and not rows[c:].any()
            assert np.array_equal(piled[i, L.uA:L.uA + c], ubn[i, :c]) and np.all(np.isposinf(piled[i, L.uA + c:L.uA + K]))
            assert np.all(np.isneginf(piled[i, L.lA + c:L.lA + K]))
            di = _abi.make_desc(n, [dict(m=3), dict(kind=_abi.TASK_IDENTITY)], c_in=c, n_bound_sets=1)
            Li = port.layout(di)
            one = np.zeros((1, Li.stride))
            one[0, Li.A[0]:Li.A[0] + 3 * n] = A0[i].reshape(-1); one[0, Li.b[0]:Li.b[0] + 3] = b0[i]; one[0, Li.b[1]:Li.b[1] + n] = bp[i]
            one[0, Li.l[0]:Li.l[0] + n] = lo[i]; one[0, Li.u[0]:Li.u[0] + n] = hi[i]
            if c:
                one[0, Li.Cin:Li.Cin + c * n] = Cn[i, :c].reshape(-1); one[0, Li.lA:Li.lA + c] = lbn[i, :c]; one[0, Li.uA:Li.uA + c] = ubn[i, :c]
            xi, _, sti, _ = port.port_solve(di, one, levels=True)
            assert sti[0] == 0 and np.abs(xi[0] - x[i]).max() <= 1e-9, i
            if c:
                binding += int((np.abs(Cn[i, :c] @ x[i] - ubn[i, :c]) < 1e-9).sum())
        return binding

    assert check(C, ub, lb, [3] * B) > 0                     # as constructed: three rows everywhere, some of them binding
    C2 = rng.standard_normal((B, K, n)); ub2 = 0.0005 + 0.004 * rng.random((B, K)); lb2 = -1e20 * np.ones((B, K))
    counts = torch.tensor([0, 1, 2, 3, 4, 5])
    hull.setRows(t(C2), t(ub2), t(lb2), count=counts)        # a different number of rows on every instance
    assert check(C2, ub2, lb2, counts.numpy()) > 0
    hull.setRows(t(C2[:, :2]), t(ub2[:, :2]), t(lb2[:, :2]), count=1)    # fewer rows than before: the old ones are gone
    check(C2, ub2, lb2, [1] * B)

"""CPU suite, part 1: the oracle itself. The C restatement (oracle/ihqp_port.c) is pinned
against (a) golden vectors of the UNMODIFIED reference (tests/golden/*.npz, made by
tests/golden/make_golden.py from oracle/_ref) and (b) the reference's own known-answer tests;
where oracle/_ref is present (this container) it is also compared live on fresh seeds."""
import numpy as np
import pytest

import common
from opensot_b200 import _abi


@pytest.mark.parametrize("tag", common.golden_tags())
def test_port_matches_reference_golden_vectors(port, tag):
    d, z = common.load_golden(tag)
    blobs = z["blobs"]
    x, xl, st, it = port.port_solve(d, blobs, levels=True)
    assert (st == 0).all()
    stable = common.stable_mask(z, x)
    err = np.abs(x - z["x_qpoases"]).max(axis=1)
    assert stable.sum() >= 0.9 * len(stable)
    assert err[stable].max() <= 1e-8, err[stable].max()
    # where eiQuadProg also solved the instance the restatement agrees with it everywhere,
    # including the instances on which qpOASES is the outlier
    ok_ei = z["ok_eiquadprog"]
    assert np.abs(x - z["x_eiquadprog"])[ok_ei].max() <= 1e-8
    L = port.layout(d)
    assert common.max_violation(d, L, blobs, x).max() <= 1e-10


def test_port_reproduces_reference_known_answers(port):
    for k in common.load_kats():
        x, st, _ = port.port_qp(k["H"], k["g"], k["C"], k["lC"], k["uC"], k["l"], k["u"],
                                eps=k["eps_factor"] * _abi.OSB_QPOASES_BASE_EPS)
        assert st == 0, k["name"]
        assert np.abs(x - np.asarray(k["expect"])).max() <= k["tol"], (k["name"], x)
        assert np.abs(x - np.asarray(k["x_reference"])).max() <= 1e-6, (k["name"], x)


def test_cost_function_is_AtWA(port):
    """tests/solvers/TestiHQP.cpp:73-146 asserts H == A'A and g == -A'b for the level cost."""
    rng = np.random.default_rng(0)
    import ctypes as C
    for m, n, wk in [(3, 7, 0), (12, 29, 1), (6, 6, 2), (1, 5, 0)]:
        A = rng.standard_normal((m, n)); b = rng.standard_normal(m); c = rng.standard_normal(n)
        if wk == 0:
            W, Wm = None, np.eye(m)
        elif wk == 1:
            W = rng.uniform(0.1, 2.0, m); Wm = np.diag(W)
        else:
            G = rng.standard_normal((m, m)); Wm = G @ G.T + m * np.eye(m); W = Wm.copy()
        H = np.zeros((n, n)); g = np.zeros(n)
        port.lib().port_cost_function(m, n, np.ascontiguousarray(A), b, wk,
                                      None if W is None else np.ascontiguousarray(W).ctypes.data,
                                      c.ctypes.data, H, g)
        assert np.allclose(H, A.T @ Wm @ A, rtol=1e-13, atol=1e-13)
        assert np.allclose(g, -A.T @ Wm @ b + c, rtol=1e-13, atol=1e-13)
        assert np.array_equal(H, H.T)


def test_single_qp_against_kkt(port):
    """random strictly convex QPs: the returned point satisfies the KKT conditions."""
    rng = np.random.default_rng(1)
    for trial in range(50):
        n, nc = int(rng.integers(2, 12)), int(rng.integers(0, 8))
        G = rng.standard_normal((n + 2, n)); H = G.T @ G; g = rng.standard_normal(n)
        Cm = rng.standard_normal((nc, n)); lC = -rng.uniform(0.1, 1, nc); uC = rng.uniform(0.1, 1, nc)
        l = -rng.uniform(0.05, 1, n); u = rng.uniform(0.05, 1, n)
        x, st, it = port.port_qp(H, g, Cm if nc else None, lC if nc else None, uC if nc else None, l, u, eps=1e-10)
        assert st == 0
        assert (x >= l - 1e-9).all() and (x <= u + 1e-9).all()
        if nc:
            cx = Cm @ x
            assert (cx >= lC - 1e-9).all() and (cx <= uC + 1e-9).all()
        # stationarity: gradient is a cone combination of active normals -> project it out
        grad = H @ x + g
        act = [np.eye(n)[i] * (1 if abs(x[i] - l[i]) < 1e-7 else -1) for i in range(n)
               if min(abs(x[i] - l[i]), abs(x[i] - u[i])) < 1e-7]
        if nc:
            cx = Cm @ x
            act += [Cm[r] * (1 if abs(cx[r] - lC[r]) < 1e-7 else -1) for r in range(nc)
                    if min(abs(cx[r] - lC[r]), abs(cx[r] - uC[r])) < 1e-7]
        if act:
            Na = np.array(act).T
            lam, *_ = np.linalg.lstsq(Na, grad, rcond=None)
            assert np.abs(Na @ lam - grad).max() <= 1e-6
            assert (lam >= -1e-6).all()
        else:
            assert np.abs(grad).max() <= 1e-6


def test_edge_cases_of_the_cascade(port):
    rng = np.random.default_rng(2)
    n = 6
    # (1) inactive task (A = 0, b unchanged: Task.h:382-386) at level 0 leaves level 1 alone
    d = _abi.make_desc(n, [dict(m=2), dict(m=n, kind=_abi.TASK_IDENTITY)], n_bound_sets=1)
    L = port.layout(d)
    blob = np.zeros(L.stride)
    blob[L.b[0]:L.b[0] + 2] = [1.0, -1.0]
    b1 = rng.uniform(-0.5, 0.5, n)
    blob[L.b[1]:L.b[1] + n] = b1
    blob[L.l[0]:L.l[0] + n] = -0.3; blob[L.u[0]:L.u[0] + n] = 0.3
    x, st, _ = port.port_solve(d, blob[None])
    assert st[0] == 0 and np.allclose(x[0], np.clip(b1, -0.3, 0.3), atol=1e-9)
    # (2) no bounds, no constraints: pure least squares cascade (pinv semantics)
    d = _abi.make_desc(n, [dict(m=2), dict(m=n, kind=_abi.TASK_IDENTITY)], n_bound_sets=0, eps_factor=1.0)
    L = port.layout(d)
    blob = np.zeros(L.stride)
    A0 = rng.standard_normal((2, n)); b0 = rng.standard_normal(2); b1 = rng.standard_normal(n)
    blob[L.A[0]:L.A[0] + 2 * n] = A0.ravel(); blob[L.b[0]:L.b[0] + 2] = b0; blob[L.b[1]:L.b[1] + n] = b1
    x, st, _ = port.port_solve(d, blob[None])
    P0 = np.linalg.pinv(A0)
    x_expect = P0 @ b0 + (np.eye(n) - P0 @ A0) @ b1
    assert st[0] == 0 and np.allclose(x[0], x_expect, atol=1e-8)
    # (3) redundant (duplicated) equality rows are accepted, inconsistent ones are reported
    d = _abi.make_desc(n, [dict(m=n, kind=_abi.TASK_IDENTITY)], c_eq=3, n_bound_sets=0)
    L = port.layout(d)
    blob = np.zeros(L.stride)
    row = rng.standard_normal(n)
    Ceq = np.vstack([row, 2 * row, rng.standard_normal(n)])
    blob[L.Ceq:L.Ceq + 3 * n] = Ceq.ravel(); blob[L.deq:L.deq + 3] = [1.0, 2.0, 0.5]
    x, st, _ = port.port_solve(d, blob[None])
    assert st[0] == 0 and np.allclose(Ceq @ x[0], [1.0, 2.0, 0.5], atol=1e-9)
    blob[L.deq + 1] = 3.0
    x, st, _ = port.port_solve(d, blob[None])
    assert (st[0] & 0xff) == _abi.INFEASIBLE
    # (4) infeasible inequality rows: x0 <= -1 and x0 >= +1 (as a row)
    d = _abi.make_desc(n, [dict(m=n, kind=_abi.TASK_IDENTITY)], c_in=1, n_bound_sets=1)
    L = port.layout(d)
    blob = np.zeros(L.stride)
    blob[L.Cin] = 1.0; blob[L.lA] = 1.0; blob[L.uA] = 2.0
    blob[L.l[0]:L.l[0] + n] = -2.0; blob[L.u[0]:L.u[0] + n] = 2.0; blob[L.u[0]] = -1.0
    x, st, _ = port.port_solve(d, blob[None])
    assert (st[0] & 0xff) == _abi.INFEASIBLE


def test_prototype_of_the_kernel_formulation_matches_port(port):
    """tests/proto_nullspace.py is the executable spec of the CUDA kernel's maths."""
    import proto_nullspace as PR
    for name, tight in [("C29", 0.3), ("C35f", 0.2), ("C38c", 0.2), ("C35h", 0.2), ("C7c", 0.4)]:
        d = port.shape(name); L = port.layout(d)
        blobs = port.synth(d, 99, 0, 6, tight)
        xp, st, _ = port.port_solve(d, blobs)
        for i in range(blobs.shape[0]):
            x, s, _ = PR.solve_stack(d, L, blobs[i])
            assert s == 0 and st[i] == 0
            assert np.abs(x - xp[i]).max() <= 1e-8


def test_live_reference_agrees_with_port_on_fresh_seeds(port):
    if not port.have_ref():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    d = port.shape("C29")
    blobs = port.synth(d, 4242, 0, 96, 0.15)
    ok, x_ref, _ = port.ref_solve(d, blobs, "qpoases")
    ok_ei, x_ei, _ = port.ref_solve(d, blobs, "eiquadprog")
    x, st, _ = port.port_solve(d, blobs)
    assert ok.all() and (st == 0).all()
    stable = np.where(ok_ei, np.abs(x_ref - x_ei).max(axis=1) <= common.STABLE_TOL,
                      np.abs(x_ref - x).max(axis=1) <= common.STABLE_TOL)
    assert stable.mean() >= 0.95
    assert np.abs(x - x_ref).max(axis=1)[stable].max() <= 1e-8

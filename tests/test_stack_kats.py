"""CPU suite: known-answer tests of the reference that exercise the STACK path (task + bounds through
AutoStack / iHQP), restated as blobs: the restatement, both kernels' logic on the warp emulator and
-- where oracle/_ref is present -- the live reference must all meet the assertion the reference's
own test makes. (tests/golden/kat_backend.json holds the single-QP tests of TestQPOases.cpp.)
The GPU twin is tests/test_gpu_parity.py::test_reference_stack_kats."""
import numpy as np
import pytest

import common
from opensot_b200 import _abi


def kat_generic_task():
    """tests/tasks/TestGenericTask.cpp:175-203 testGenericTaskWithQPOASES: A = [1 1], b = 1, bounds
    +-5, iHQP(stack, bounds, eps = 1.): EXPECT_NEAR(x.sum(), b[0], 1e-9)"""
    d = _abi.make_desc(2, [dict(m=1)], n_bound_sets=1, eps_factor=1.0)

    def fill(L, blob):
        blob[L.A[0]:L.A[0] + 2] = [1.0, 1.0]; blob[L.b[0]] = 1.0
        blob[L.l[0]:L.l[0] + 2] = -5.0; blob[L.u[0]:L.u[0] + 2] = 5.0
    return d, fill, lambda x, blob, L: abs(x.sum() - 1.0) <= 1e-9


def kat_pinv_vs_qp(seed):
    """tests/solvers/TestBasicAlgebra.cpp:46-108 testPinvVSQP: A 12 x 30 and b uniform in [-1, 1]
    (Eigen setRandom), bounds +-1, H = A'A, g = -A'b: EXPECT_NEAR((A x)[i], b[i], 1e-5)"""
    rng = np.random.default_rng(seed)
    A, b = rng.uniform(-1, 1, (12, 30)), rng.uniform(-1, 1, 12)
    d = _abi.make_desc(30, [dict(m=12)], n_bound_sets=1, eps_factor=2e2)

    def fill(L, blob):
        blob[L.A[0]:L.A[0] + 360] = A.reshape(-1); blob[L.b[0]:L.b[0] + 12] = b
        blob[L.l[0]:L.l[0] + 30] = -1.0; blob[L.u[0]:L.u[0] + 30] = 1.0
    # 18 of the 30 directions are fixed by eps |x|^2 alone: H = A'A + eps I has condition ~1e11, and
    # solvers that factor H (qpOASES, the restatement) land ~1e-5 from the exact minimiser
    # x* = A'(AA' + eps I)^-1 b, which the kernels (LQ of A, no H) reproduce to round-off
    xstar = A.T @ np.linalg.solve(A @ A.T + d.eps * np.eye(12), b)
    assert np.abs(xstar).max() < 1.0
    return d, fill, lambda x, blob, L: np.abs(A @ x - b).max() <= 1e-5, dict(exact=xstar, xtol=1e-4)


def kat_simple_qp(seed):
    """examples/cpp/simple_qp.cpp: x in R^5, cost mytask1 + mytask2 = ||x - 2||^2 + ||R x - r||^2 (R, r
    uniform in [-1, 1]), subject to -1 <= sum(x) <= 0.3, iHQP(stack, 2e2, qpOASES). The example
    prints its solution; what is checkable is the constraint and the optimality of the answer:
    the gradient is a non-positive multiple of the constraint normal (upper side active) or zero."""
    rng = np.random.default_rng(seed)
    R, r = rng.uniform(-1, 1, (5, 5)), rng.uniform(-1, 1, 5)
    A = np.vstack([np.eye(5), R]); b = np.r_[2.0 * np.ones(5), r]
    d = _abi.make_desc(5, [dict(m=10)], c_in=1, n_bound_sets=0, eps_factor=2e2)

    def fill(L, blob):
        blob[L.A[0]:L.A[0] + 50] = A.reshape(-1); blob[L.b[0]:L.b[0] + 10] = b
        blob[L.Cin:L.Cin + 5] = 1.0; blob[L.lA] = -1.0; blob[L.uA] = 0.3

    def check(x, blob, L):
        s = x.sum()
        grad = A.T @ (A @ x - b) + d.eps * x
        if not (-1.0 - 1e-9 <= s <= 0.3 + 1e-9):
            return False
        if abs(s - 0.3) <= 1e-9:         # upper side active: grad = -mu * 1, mu >= 0
            return np.abs(grad - grad.mean()).max() <= 1e-8 and grad.mean() <= 1e-8
        return np.abs(grad).max() <= 1e-8
    return d, fill, check


CASES = [("generic_task", kat_generic_task())] + [("pinv_vs_qp_%d" % s, kat_pinv_vs_qp(s)) for s in (1, 2, 3)] \
    + [("simple_qp_%d" % s, kat_simple_qp(s)) for s in (1, 2)]


def build(port_or_layout, d, fill):
    L = port_or_layout(d)
    blob = np.zeros((1, L.stride))
    fill(L, blob[0])
    return L, blob


@pytest.mark.parametrize("name,case", CASES, ids=[c[0] for c in CASES])
def test_restatement_emulated_kernels_and_reference_meet_the_reference_assertions(port, name, case):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import emulib
    d, fill, check = case[:3]
    extra = case[3] if len(case) > 3 else {}
    L, blob = build(port.layout, d, fill)
    x, st, _ = port.port_solve(d, blob)
    assert st[0] == 0 and check(x[0], blob[0], L), "restatement"
    answers = [x[0]]
    for variant in (emulib.GENERIC_F64, emulib.RR_F64):
        r = emulib.solve(d, blob, variant=variant)
        if r is None:                       # shape outside the register-row instances
            assert variant == emulib.RR_F64
            continue
        assert r[1][0] == 0 and check(r[0][0], blob[0], L), "emulated kernel %d" % variant
        if "exact" in extra:
            assert np.abs(r[0][0] - extra["exact"]).max() <= 1e-9, "kernel vs the exact minimiser"
        answers.append(r[0][0])
    if port.have_ref():
        ok, xr, _ = port.ref_solve(d, blob)
        assert ok[0] and check(xr[0], blob[0], L), "live reference"
        answers.append(xr[0])
    for a in answers[1:]:
        assert np.abs(a - answers[0]).max() <= extra.get("xtol", common.PARITY_TOL)


def test_third_oracle_hcod_arbitrates(port):
    """OpenSoT::solvers::HCOD over the vendored soth (a true lexicographic solver: no eps, another
    algorithm family; SURVEY 8c "third oracle", built by `make -C oracle hcod`) as arbiter:
    - on well-conditioned C29 stacks it agrees with iHQP+qpOASES and with the restatement to 1e-9;
    - on the reference's 12 x 30 test, where qpOASES and the restatement sit ~1e-5 from the exact
      minimiser, HCOD sides with the kernels (both within 1e-9 of it)."""
    if not (port.have_ref() and port.have_hcod()):
        pytest.skip("oracle/_ref/ref_driver_hcod not built")
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import emulib
    d = port.shape("C29")
    blobs = port.synth(d, 5, 0, 12, 0.05)
    okh, xh, _ = port.ref_solve(d, blobs, backend="hcod")
    okq, xq, _ = port.ref_solve(d, blobs)
    xp, stp, _ = port.port_solve(d, blobs)
    assert okh.all() and okq.all() and (stp == 0).all()
    assert np.abs(xh - xq).max() <= 1e-9 and np.abs(xh - xp).max() <= 1e-9
    xe = emulib.solve(d, blobs, variant=emulib.RR_F64)[0]
    assert np.abs(xe - xh).max() <= 1e-9
    for seed in (1, 2, 3):
        dd, fill, check, extra = kat_pinv_vs_qp(seed)
        L, blob = build(port.layout, dd, fill)
        ok, x, _ = port.ref_solve(dd, blob, backend="hcod")
        assert ok[0] and check(x[0], blob[0], L)
        assert np.abs(x[0] - extra["exact"]).max() <= 1e-9
        xk = emulib.solve(dd, blob, variant=emulib.RR_F64)[0]
        assert np.abs(xk[0] - x[0]).max() <= 1e-9
        _, xq, _ = port.ref_solve(dd, blob)
        assert np.abs(xq[0] - x[0]).max() > 1e-7          # qpOASES is the outlier on eps-determined directions

"""CPU suite: logging parity (opensot_b200/qplog.py). The logged
per-level QPs -- H_i, g_i, A_i, lA_i, uA_i, l_i, u_i under the reference's names
(src/solvers/BackEnd.cpp:111-130) -- must be the QPs the reference's cascade solves: handing each
of them (with the back-end's eps on the diagonal, QPOasesBackEnd.cpp:253-255) to the oracle's
single-QP solver reproduces solution_i, on every level path the blob layout has."""
import numpy as np
import pytest

import common
from opensot_b200 import _abi, qplog


@pytest.mark.parametrize("name,count", [("C29", 6), ("C35f", 4), ("C29w", 4), ("C29r", 4), ("C29c", 4), ("C16e", 6), ("C20d", 4)])
def test_logged_level_qps_reproduce_the_level_solutions(port, name, count):
    d = port.shape(name)
    L = port.layout(d)
    blobs = port.synth(d, 808, 0, count, 0.2)
    x, xl, st, _ = port.port_solve(d, blobs, levels=True)
    log = qplog.level_qps(d, L, blobs, xl)
    n = d.n
    checked = 0
    for i in np.flatnonzero(st == 0):
        for k in range(d.n_levels):
            H, g = log["H_%d" % k][i], log["g_%d" % k][i]
            assert np.allclose(H, H.T, atol=1e-12)
            A = log.get("A_%d" % k)
            xk, stk, _ = port.port_qp(H, g, None if A is None else A[i],
                                      None if A is None else log["lA_%d" % k][i],
                                      None if A is None else log["uA_%d" % k][i],
                                      log["l_%d" % k][i] if d.n_bound_sets else None,
                                      log["u_%d" % k][i] if d.n_bound_sets else None, eps=d.eps)
            if stk != 0:       # dependent equality rows in a Goldfarb-Idnani QP: no verdict
                continue
            assert np.abs(xk - log["solution_%d" % k][i]).max() <= common.PARITY_TOL, (name, i, k)
            checked += 1
    assert checked >= count
    assert log["solution_%d" % (d.n_levels - 1)].shape == (count, n)
    assert np.array_equal(log["solution_%d" % (d.n_levels - 1)][st == 0], x[st == 0])


def test_names_shapes_and_files(tmp_path, port):
    d = port.shape("C35f")
    L = port.layout(d)
    blobs = port.synth(d, 3, 0, 2, 0.1)
    _, xl, _, _ = port.port_solve(d, blobs, levels=True)
    log = qplog.level_qps(d, L, blobs, xl)
    n, ce = d.n, d.c_eq
    assert log["A_0"].shape == (2, ce, n) and log["A_1"].shape == (2, ce + d.m[0], n)
    assert log["A_2"].shape == (2, ce + d.m[0] + d.m[1], n)
    assert np.array_equal(log["lA_0"], log["uA_0"])                 # Ceq x = deq as lA = uA rows
    assert np.allclose(log["lA_2"][:, ce:ce + d.m[0]], np.einsum("bij,bj->bi", log["A_2"][:, ce:ce + d.m[0]], xl[:, 0]))
    tl = qplog.task_logs(d, L, blobs, ids=["com", "wrists", "postural"])
    assert tl["com_A"].shape == (2, d.m[0], n) and "wrists_W" in tl and tl["postural_A"].shape == (2, n, n)
    qplog.save(str(tmp_path / "log.npz"), {**log, **tl})
    z = np.load(tmp_path / "log.npz")
    assert np.array_equal(z["H_1"], log["H_1"])
    qplog.save(str(tmp_path / "log.mat"), {**log, **tl})
    from scipy.io import loadmat
    assert np.allclose(loadmat(str(tmp_path / "log.mat"))["g_2"], log["g_2"])
    # a switched-off level contributes neither a QP nor optimality rows
    _, xl2, _, _ = port.port_solve(d, blobs, active=[True, False, True], levels=True)
    log2 = qplog.level_qps(d, L, blobs, xl2, active=[True, False, True])
    assert "H_1" not in log2 and log2["A_2"].shape == (2, ce + d.m[0], n)

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

"""Shared helpers of the test-suite: golden fixtures and comparison rules."""
import ctypes as C
import glob
import json
import os

import numpy as np

from opensot_b200 import _abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# joint-velocity infinity-norm parity bound stated by BASELINE.json north_star
PARITY_TOL = 1e-6
# two reference back-ends / the restatement agree this well on a well-defined instance
STABLE_TOL = 1e-8


def golden_tags():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_golden(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    d = _abi.StackDesc.from_buffer_copy(z["desc"].tobytes())
    return d, z


def stable_mask(z, x_port):
    """Instances on which the reference answer is well-defined:
    qpOASES succeeded and agrees to 1e-8 with eiQuadProg where that succeeded, else with the
    restatement. The remainder is counted and reported, never tuned to."""
    qp, ei = z["x_qpoases"], z["x_eiquadprog"]
    ok_qp, ok_ei = z["ok_qpoases"], z["ok_eiquadprog"]
    d_ei = np.abs(qp - ei).max(axis=1)
    d_po = np.abs(qp - x_port).max(axis=1)
    return ok_qp & np.where(ok_ei, d_ei <= STABLE_TOL, d_po <= STABLE_TOL)


def load_kats():
    with open(os.path.join(GOLDEN, "kat_backend.json")) as f:
        return json.load(f)


def kat_to_stack(k, layout_fn):
    """BackEnd-level QP -> single-level OSB_TASK_HESSIAN stack blob."""
    n = len(k["g"])
    nc = 0 if k["C"] is None else np.asarray(k["C"]).reshape(-1, n).shape[0]
    d = _abi.make_desc(n, [dict(m=0, kind=_abi.TASK_HESSIAN)], c_in=nc, n_bound_sets=1,
                       eps_factor=k["eps_factor"])
    L = layout_fn(d)
    blob = np.zeros(L.stride)
    blob[L.A[0]:L.A[0] + n * n] = np.asarray(k["H"], dtype=float).reshape(-1)
    blob[L.b[0]:L.b[0] + n] = k["g"]
    if nc:
        blob[L.Cin:L.Cin + nc * n] = np.asarray(k["C"], dtype=float).reshape(-1)
        blob[L.lA:L.lA + nc] = k["lC"]
        blob[L.uA:L.uA + nc] = k["uC"]
    blob[L.l[0]:L.l[0] + n] = k["l"]
    blob[L.u[0]:L.u[0] + n] = k["u"]
    return d, blob


from oracle.portlib import bounds_of, max_violation, live_stable_mask, REF_FEAS_TOL  # noqa: E402,F401

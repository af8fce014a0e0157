"""Generates the golden vectors under tests/golden/ from the UNMODIFIED reference compiled
by oracle/Makefile (oracle/_ref/ref_driver: AutoStack::update() + iHQP::solve(), qpOASES and
eiQuadProg back-ends through the BackEndFactory dlopen boundary). Needs /root/reference (to
build oracle/_ref); run from the repo root:   python tests/golden/make_golden.py
The .npz files are committed; nothing at test time needs /root/reference."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import portlib as P  # noqa: E402
from opensot_b200 import _abi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (file tag, shape, eps factor, seed, count, tight)
STACK_CASES = [
    ("C29_t10", "C29", 2e2, 11, 48, 0.10),
    ("C29_t40", "C29", 2e2, 12, 48, 0.40),
    ("C29_eps1e6", "C29", 1e6, 13, 24, 0.20),   # eps used by the published benchmark, coman_ik.cpp:453
    ("C35f_t10", "C35f", 2e2, 14, 24, 0.10),
    ("C38c_t10", "C38c", 2e2, 15, 24, 0.10),
    ("C35h_t10", "C35h", 2e2, 16, 24, 0.10),
    ("C12_t30", "C12", 2e2, 17, 64, 0.30),
    ("C7c_t30", "C7c", 2e2, 18, 64, 0.30),
]


def stack_case(tag, shape, eps_factor, seed, count, tight):
    d = P.shape(shape, eps_factor)
    blobs = P.synth(d, seed, 0, count, tight)
    ok_qp, x_qp, xl_qp = P.ref_solve(d, blobs, "qpoases")
    ok_ei, x_ei, xl_ei = P.ref_solve(d, blobs, "eiquadprog")
    np.savez_compressed(
        os.path.join(OUT, tag + ".npz"), desc=np.frombuffer(bytes(d), dtype=np.uint8),
        shape=shape, eps_factor=eps_factor, seed=seed, tight=tight, blobs=blobs,
        ok_qpoases=ok_qp, x_qpoases=x_qp, xlev_qpoases=xl_qp,
        ok_eiquadprog=ok_ei, x_eiquadprog=x_ei)
    dq = np.abs(x_qp - x_ei).max(axis=1)
    print("%-12s qpOASES ok %d/%d, eiQuadProg ok %d/%d, max|qp-ei| on ei-ok %.2e"
          % (tag, ok_qp.sum(), count, ok_ei.sum(), count, dq[ok_ei].max() if ok_ei.any() else -1))


# BackEnd-level known answers restated from the reference's own tests. Each is solved here
# through the compiled reference (single-level OSB_TASK_HESSIAN stack -> BackEnd) so that the
# file holds both the expectation written in the reference test and the reference's output.
def kat_cases():
    K = []
    I2 = np.eye(2)
    K.append(dict(name="testSimpleProblem", cite="tests/solvers/TestQPOases.cpp:346-377 (fixture :86-110)",
                  eps_factor=1e-9, H=I2, g=[-5.0, 5.0], C=np.zeros((2, 2)), lC=[-10, -10], uC=[10, 10],
                  l=[-10, -10], u=[10, 10], expect=[5.0, -5.0], tol=1e-14))
    T = np.array([[1.0, 1, 1]]); b = np.array([10.0])
    K.append(dict(name="test_update_constraint_1", cite="tests/solvers/TestQPOases.cpp:208-239",
                  eps_factor=1e4, H=T.T @ T, g=-T.T @ b, C=np.zeros((1, 3)), lC=[0], uC=[0],
                  l=[-10] * 3, u=[10] * 3, expect=[3.333] * 3, tol=1e-3))
    K.append(dict(name="test_update_constraint_2", cite="tests/solvers/TestQPOases.cpp:241-253",
                  eps_factor=1e4, H=T.T @ T, g=-T.T @ b, C=np.array([[1.0, 0, 1]]), lC=[20], uC=[20],
                  l=[-10] * 3, u=[10] * 3, expect=[10.0, -10.0, 10.0], tol=1e-6))
    T = np.array([[1.0, 1, 1], [0, 1, 1]]); b = np.array([6.0, 5])
    K.append(dict(name="test_update_task_1", cite="tests/solvers/TestQPOases.cpp:274-303",
                  eps_factor=1.0, H=T.T @ T, g=-T.T @ b, C=None, lC=None, uC=None,
                  l=[-10] * 3, u=[10] * 3, expect=[1.0, 2.5, 2.5], tol=1e-6))
    T = np.array([[1.0, 1, 1], [0, 1, 1], [1, 1, 0]]); b = np.array([6.0, 5, 3])
    K.append(dict(name="test_update_task_2", cite="tests/solvers/TestQPOases.cpp:305-319",
                  eps_factor=1.0, H=T.T @ T, g=-T.T @ b, C=None, lC=None, uC=None,
                  l=[-10] * 3, u=[10] * 3, expect=[1.0, 2.0, 3.0], tol=1e-6))
    T = np.array([[1.0, 1, 1], [0, 1, 1], [1, 1, 0], [1, 0, 1]]); b = np.array([6.0, 5, 3, 3])
    K.append(dict(name="test_update_task_3", cite="tests/solvers/TestQPOases.cpp:321-339",
                  eps_factor=1.0, H=T.T @ T, g=-T.T @ b, C=None, lC=None, uC=None,
                  l=[-10] * 3, u=[10] * 3, expect=[0.5714, 2.5714, 2.5714], tol=1e-4))
    return K


def kat_to_blob(k):
    """BackEnd-level QP -> single-level OSB_TASK_HESSIAN stack (rows with lC==uC stay in Cin)."""
    n = len(k["g"])
    nc = 0 if k["C"] is None else np.asarray(k["C"]).shape[0]
    d = _abi.make_desc(n, [dict(m=0, kind=_abi.TASK_HESSIAN)], c_in=nc, n_bound_sets=1,
                       eps_factor=k["eps_factor"])
    L = P.layout(d)
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


def main():
    if not P.have_ref():
        raise SystemExit("oracle/_ref missing: run `make -C oracle ref` (needs /root/reference)")
    for c in STACK_CASES:
        stack_case(*c)
    out = []
    for k in kat_cases():
        x_ref = P.ref_qp(k)
        err = np.abs(x_ref - np.asarray(k["expect"])).max()
        print("KAT %-26s reference x = %s  |x-expect| = %.2e (tol %g)" % (k["name"], np.array2string(x_ref, precision=6), err, k["tol"]))
        assert err <= k["tol"], "reference build disagrees with its own test expectation"
        j = {a: (np.asarray(v).tolist() if isinstance(v, (np.ndarray, list)) else v) for a, v in k.items()}
        j["x_reference"] = x_ref.tolist()
        out.append(j)
    with open(os.path.join(OUT, "kat_backend.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

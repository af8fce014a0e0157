#!/usr/bin/env python
"""Small batch through every kernel family, checked against the oracle restatement; meant to be run
under compute-sanitizer on the GPU box (tools/evidence_r02.sh):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Families: register-row kernel 30x1 / 30x2 / 38x2 (fp64, fp32), the CTA-per-stack kernel where the
engine picks it, the generic kernel (fp64, fp32), sot_gather_kernel, sot_posture_kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import opensot_b200 as ob
from opensot_b200 import synth, _abi
from oracle import portlib as P


def main():
    B = int(os.environ.get("SAN_BATCH", "24"))
    dev = torch.device("cuda", 0)
    worst = 0.0
    for name in ("C29", "C35f", "C38c", "C35h", "C12", "C64c", "C29w", "C29r"):
        for prec in ("f64", "f32"):
            if prec == "f32" and name in ("C29w", "C29r"):
                continue    # explicit reduced Hessians of rank-deficient levels: fp64 engine only (OSB_E_PRECISION)
            desc = synth.shape(name, 2e2 if prec == "f64" else _abi.F32_EPS_FACTOR)
            eng = ob.BatchedStackSolver(desc, B, device=0, precision=prec, store_levels=True)
            b64 = synth.fill(desc, 5, 0, B, 0.2)
            x, st, xl = eng.solve_host(eng.pack(b64))
            xp, stp, _ = P.port_solve(desc, b64)
            ok = (st == 0) & (stp == 0)
            err = float(np.abs(x[ok] - xp[ok]).max()) if ok.any() else float("nan")
            tol = 1e-6 if prec == "f64" else 5e-3
            print("%-5s %s %-28s solved %d/%d  max|x-x_oracle| %.2e" % (name, prec, eng.kernel_name(), int(ok.sum()), B, err), flush=True)
            assert ok.sum() >= B - 2 and err <= tol, (name, prec, err)
            if prec == "f64":
                worst = max(worst, err)
            eng.close()
    # device-side piling (sot_gather_kernel) and joint-state producers (sot_posture_kernel)
    desc = synth.shape("C29")
    L = ob.blob_layout(desc)
    eng = ob.BatchedStackSolver(desc, B, device=0)
    blobs = torch.from_numpy(synth.fill(desc, 9, 0, B, 0.2)).to(dev)
    n = desc.n
    A1 = blobs[:, L.A[1]:L.A[1] + 12 * n].reshape(B, 12, n).contiguous()
    out = blobs.clone()
    out[:, L.A[1]:L.A[1] + 12 * n] = 0
    eng.gather_device([dict(src=A1[:, :6].contiguous(), rows=6, cols=n, dst_off=L.A[1]),
                       dict(src=A1[:, 6:].contiguous(), rows=6, cols=n, dst_off=L.A[1] + 6 * n)], out)
    torch.cuda.synchronize()
    assert torch.equal(out, blobs), "piling differs"
    q = torch.zeros((B, n), dtype=torch.float64, device=dev)
    lim = torch.ones(n, dtype=torch.float64, device=dev)
    eng.produce_posture_device(out, q, q_ref=q + 0.01, q_min=-lim, q_max=lim, qdot_max=lim, lam=1.0, dT=0.01,
                               postural_level=2, joint_limit_set=0, velocity_limit_set=1)
    torch.cuda.synchronize()
    got = out[:, L.b[2]:L.b[2] + n]
    assert torch.allclose(got, torch.full_like(got, 0.01)), "posture producer differs"
    # Cartesian / CoM producers and the state-input path (sot_kinematics_kernel, sot_template_kernel)
    from opensot_b200 import kinematics as K
    tree = K.load_npz(os.path.join(ROOT, "tests", "golden", "robots", "coman_tree.npz"))
    tmpl = np.zeros(L.stride)
    tmpl[L.w[1]:L.w[1] + 12] = [0.1] * 6 + [1.0] * 6
    eng.kinematics_create(tree)
    pose = lambda p: np.concatenate([np.eye(3).reshape(-1), p])   # noqa: E731
    eng.states_configure(tmpl, [dict(kind="com", level=0, lam=0.1, desired=np.array([0.0, 0.0, 0.02])),
                                dict(kind="cartesian", distal="l_wrist", level=1, row0=0, lam=0.1, desired=pose([0.005, 0.156, -0.05])),
                                dict(kind="cartesian", distal="r_wrist", base="Waist", level=1, row0=6, lam=0.1, desired=pose([0.005, -0.156, -0.05]))],
                         q_ref=np.zeros(n), q_min=tree.q_min, q_max=tree.q_max, qdot_max=np.minimum(tree.qdot_max, 4.0), lam=0.05, dT=0.01,
                         postural_level=2, joint_limit_set=0, velocity_limit_set=1)
    rng = np.random.default_rng(3)
    xs, sts = eng.solve_states_host(0.12 * rng.uniform(tree.q_min, tree.q_max, size=(B, n)))
    assert (sts == 0).all(), sts
    print("gather + posture + kinematics kernels and the state-input path ok; worst fp64 parity %.2e" % worst)


if __name__ == "__main__":
    main()

"""Logging parity for drop-in debugging -- host only, off the hot path.

The reference dumps, per priority level i, the QP its back-end was given:
`BackEnd::log` adds H_i, g_i, A_i, lA_i, uA_i, l_i, u_i, solution_i (src/solvers/BackEnd.cpp:111-130),
`iHQP::_log` adds Hr, gr of a user regularisation task and calls the back-ends
(src/solvers/iHQP.cpp:403-415), `Task::log` adds <id>_A, <id>_b, <id>_W, <id>_c
(include/OpenSoT/Task.h:437-453), all into a MatLogger2 .mat file. The batched engine never
forms those matrices, so this module rebuilds them on the host, from the
instance blobs and the per-level solutions (`store_levels=True`), under the same names and with
a leading batch axis -- the file a user of the reference would diff against their own dump:

    H_i      = A_i' W_i A_i (+ Hr)                 iHQP.cpp:129-162, 274-278   (eps is added inside
                                                    the back-end's solve(), QPOasesBackEnd.cpp:253-255:
                                                    logged separately as eps_regularisation)
    g_i      = -A_i' W_i b_i + c_i (+ gr)
    A_i      = [Cloc_i ; Ceq ; Cin ; A_0 ; ... ; A_{i-1}]   the level's own rows, the global rows
    lA_i,uA_i= [lCl_i ; deq ; lA ; A_j x_j* ...]      (iHQP.cpp:190-198: the task's constraints, then
                                                    the stack's), equalities as lA = uA rows
                                                    (constraints/Aggregated.cpp:159-172), then one
                                                    optimality block per earlier ACTIVE level
                                                    (iHQP.cpp:164-170, 285-332; the reference keeps
                                                    -1 <= 0 x <= 1 placeholder rows for switched-off
                                                    levels, iHQP.cpp:303-309: omitted here)
    l_i, u_i = max / min over the bound sets        constraints/Aggregated.cpp:141-148
    solution_i = x_i*

Nothing here is used by the solve path; numpy only.
"""
import numpy as np

from . import _abi


def _level_task(desc, L, blobs, k):
    """(A [B,m,n], b [B,m], W [B?,m,m] or None, c [B,n] or None) of level k; HESSIAN levels return (None, ...)."""
    B, n, m = blobs.shape[0], desc.n, desc.m[k]
    kind, wk = desc.task_kind[k], desc.weight_kind[k]
    if kind == _abi.TASK_IDENTITY:
        A = np.broadcast_to(np.eye(n), (B, n, n))
    else:
        A = blobs[:, L.A[k]:L.A[k] + m * n].reshape(B, m, n)
    b = blobs[:, L.b[k]:L.b[k] + m]
    W = None
    if wk == _abi.W_DIAG:
        w = blobs[:, L.w[k]:L.w[k] + m]
        W = np.einsum("bi,ij->bij", w, np.eye(m))
    elif wk == _abi.W_DENSE:
        W = blobs[:, L.w[k]:L.w[k] + m * m].reshape(B, m, m)
    c = blobs[:, L.c[k]:L.c[k] + n] if desc.has_c[k] else None
    return A, b, W, c


def level_qps(desc, layout, blobs, x_levels=None, active=None):
    """dict of arrays with the reference's log names (see the module docstring); x_levels
    [B, n_levels, n] supplies the optimality right-hand sides and solution_i (omitted if None)."""
    blobs = np.asarray(blobs, dtype=np.float64)
    d, L = desc, layout
    B, n = blobs.shape[0], d.n
    active = [True] * d.n_levels if active is None else [bool(a) for a in active]
    out = {"eps_regularisation": np.float64(d.eps)}
    lo = np.full((B, n), -_abi.OSB_INFTY); hi = np.full((B, n), _abi.OSB_INFTY)
    for s in range(d.n_bound_sets):
        lo = np.maximum(lo, blobs[:, L.l[s]:L.l[s] + n]); hi = np.minimum(hi, blobs[:, L.u[s]:L.u[s] + n])
    Hr = gr = None
    if d.m_reg > 0:
        Ar = blobs[:, L.Ar:L.Ar + d.m_reg * n].reshape(B, d.m_reg, n)
        br = blobs[:, L.br:L.br + d.m_reg]
        wr = blobs[:, L.wr:L.wr + d.m_reg] if d.reg_weight_kind == _abi.W_DIAG else np.ones((B, d.m_reg))
        Hr = np.einsum("bri,br,brj->bij", Ar, wr, Ar)
        gr = -np.einsum("bri,br->bi", Ar, wr * br)
        out["Hr"], out["gr"] = Hr, gr
    glob_A, glob_l, glob_u = [], [], []
    if d.c_eq:
        glob_A.append(blobs[:, L.Ceq:L.Ceq + d.c_eq * n].reshape(B, d.c_eq, n))
        deq = blobs[:, L.deq:L.deq + d.c_eq]
        glob_l.append(deq); glob_u.append(deq)
    if d.c_in:
        glob_A.append(blobs[:, L.Cin:L.Cin + d.c_in * n].reshape(B, d.c_in, n))
        glob_l.append(blobs[:, L.lA:L.lA + d.c_in]); glob_u.append(blobs[:, L.uA:L.uA + d.c_in])
    opt_A, opt_b = [], []          # optimality rows of the earlier active levels
    for k in range(d.n_levels):
        if not active[k]:
            continue
        if d.task_kind[k] == _abi.TASK_HESSIAN:
            H = blobs[:, L.A[k]:L.A[k] + n * n].reshape(B, n, n).copy()
            g = blobs[:, L.b[k]:L.b[k] + n].copy()
            A = None
        else:
            A, b, W, c = _level_task(d, L, blobs, k)
            WA = A if W is None else np.einsum("bij,bjk->bik", W, A)
            H = np.einsum("bji,bjk->bik", A, WA)
            Wb = b if W is None else np.einsum("bij,bj->bi", W, b)
            g = -np.einsum("bji,bj->bi", A, Wb)
            if c is not None:
                g = g + c
        if Hr is not None:
            H = H + Hr; g = g + gr
        rows_A, rows_l, rows_u = [], [], []
        if d.c_loc[k]:
            cl = d.c_loc[k]
            rows_A.append(blobs[:, L.Cl[k]:L.Cl[k] + cl * n].reshape(B, cl, n))
            rows_l.append(blobs[:, L.lCl[k]:L.lCl[k] + cl]); rows_u.append(blobs[:, L.uCl[k]:L.uCl[k] + cl])
        rows_A += glob_A + opt_A; rows_l += glob_l + opt_b; rows_u += glob_u + opt_b
        out["H_%d" % k], out["g_%d" % k] = H, g
        if rows_A:
            out["A_%d" % k] = np.concatenate(rows_A, axis=1)
            out["lA_%d" % k] = np.clip(np.concatenate(rows_l, axis=1), -_abi.OSB_INFTY, _abi.OSB_INFTY)
            out["uA_%d" % k] = np.clip(np.concatenate(rows_u, axis=1), -_abi.OSB_INFTY, _abi.OSB_INFTY)
        if d.n_bound_sets:
            out["l_%d" % k], out["u_%d" % k] = lo, hi
        if x_levels is not None:
            xk = np.asarray(x_levels, dtype=np.float64)[:, k]
            out["solution_%d" % k] = xk
            if A is not None:          # this level's optimality rows for the levels below it
                opt_A.append(np.array(A)); opt_b.append(np.einsum("bij,bj->bi", A, xk))
    return out


def task_logs(desc, layout, blobs, ids=None):
    """<id>_A, <id>_b, <id>_W, <id>_c per level, the names of Task::log (Task.h:437-453); ids default to level0.."""
    blobs = np.asarray(blobs, dtype=np.float64)
    out = {}
    for k in range(desc.n_levels):
        if desc.task_kind[k] == _abi.TASK_HESSIAN:
            continue
        tid = ids[k] if ids else "level%d" % k
        A, b, W, c = _level_task(desc, layout, blobs, k)
        out[tid + "_A"], out[tid + "_b"] = np.array(A), b
        if W is not None:
            out[tid + "_W"] = W
        if c is not None:
            out[tid + "_c"] = c
    return out


def save(path, arrays):
    """.mat (what MatLogger2 writes; needs scipy) or .npz by extension"""
    if path.endswith(".mat"):
        from scipy.io import savemat
        savemat(path, arrays)
    else:
        np.savez_compressed(path, **arrays)

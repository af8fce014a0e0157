"""numpy prototype of the kernel's formulation (null-space cascade + dual active set on
y = [x; Cin x; Cloc x]). Used on CPU to validate the MATH of opensot_b200/csrc/sot_kernel.cuh
against the oracle before GPU time is spent; mirrors the kernel step by step (same names).
Not a product path and not an oracle: a development aid kept as an executable spec."""
import numpy as np

INF = 1e20
TRACE = False
DEP_TOL = 1e-22


def clip(v):
    return np.clip(v, -INF, INF)


def sweep(ZB, z0, nz, M, tol_rank):
    """Householder sweep: returns rank, piv; modifies ZB[:, z0:z0+nz] and M in place."""
    m = M.shape[0]
    rn2 = (M[:, :nz] ** 2).sum(axis=1)
    piv = -np.ones(m, dtype=int)
    zc = 0
    for r in range(m):
        if zc >= nz:
            break
        a = M[r, zc:nz]
        nrm2 = a @ a
        if not (nrm2 > tol_rank ** 2 * rn2[r]) or nrm2 <= 0:
            M[r, zc:nz] = 0
            continue
        rho = -np.sqrt(nrm2) if a[0] > 0 else np.sqrt(nrm2)
        beta = 1.0 / (rho * (rho - a[0]))
        v = a.copy(); v[0] -= rho
        blk = ZB[:, z0 + zc:z0 + nz]
        blk -= beta * np.outer(blk @ v, v)
        rest = M[r + 1:, zc:nz]
        rest -= beta * np.outer(rest @ v, v)
        M[r, zc:nz] = 0; M[r, zc] = rho
        piv[r] = zc
        zc += 1
    return zc, piv


def gi(N, y, ylo, yhi, nz, tol, max_iter):
    """dual active set on box; N is NR x nz (modified in place), y modified in place."""
    NR = N.shape[0]
    R = np.zeros((nz, nz)); u = np.zeros(nz + 1)
    act = []  # (row, sign)
    where = -np.ones(NR, dtype=int)
    iters = 0
    parked = False
    while True:
        best, bidx, bs = 0.0, -1, 0
        for i in range(NR):
            if where[i] != -1:
                continue
            vl = (ylo[i] - y[i]) - tol * max(1.0, abs(ylo[i])) if ylo[i] > -INF else -1.0
            vh = (y[i] - yhi[i]) - tol * max(1.0, abs(yhi[i])) if yhi[i] < INF else -1.0
            cand, side = (vh, -1) if vh > vl else (vl, +1)
            if cand > best:
                best, bidx, bs = cand, i, side
        if bidx < 0:
            return 0, iters, act, where
        ip, sgn = bidx, bs
        up = 0.0
        while True:
            iters += 1
            if iters > max_iter:
                return 1, iters, act, where
            q = len(act)
            d = sgn * N[ip, :nz].copy()
            dd = d @ d; zz = d[q:] @ d[q:]
            dy = N[:, q:nz] @ d[q:]
            t1, lk = np.inf, -1
            r = np.zeros(q)
            if q > 0:
                r = np.linalg.solve(R[:q, :q], d[:q]) if q else r
                for k in range(q):
                    ak = act[k][0]
                    if r[k] > 0 and not (ylo[ak] == yhi[ak]):
                        ratio = u[k] / r[k]
                        if ratio < t1:
                            t1, lk = ratio, k
            lim = ylo[ip] if sgn > 0 else yhi[ip]
            s = sgn * (y[ip] - lim)
            indep = (zz > DEP_TOL * dd) and q < nz
            t2 = -s / zz if indep else np.inf
            t = min(t1, t2)
            if TRACE:
                print("it %d ip %d sgn %+.0f q %d nz %d s %.6e zz %.6e dd %.6e t1 %.6e lk %d t2 %.6e" % (iters, ip, sgn, q, nz, s, zz, dd, t1, lk, t2))
            if not t < np.inf:
                if -s <= 1e-8 * max(1.0, abs(lim)):
                    where[ip] = -2; parked = True
                    break
                return 2, iters, act, where
            if indep:
                y += t * dy
            u[:q] -= t * r
            up += t
            if indep and t2 <= t1:
                dq = d[q]
                rho = -np.sqrt(zz) if dq > 0 else np.sqrt(zz)
                beta = 1.0 / (rho * (rho - dq))
                v = d[q:].copy(); v[0] = dq - rho
                w = beta * (dy - rho * N[:, q])
                N[:, q:nz] -= np.outer(w, v)
                R[:q, q] = d[:q]; R[q, q] = rho
                act.append((ip, sgn)); where[ip] = q; u[q] = up
                if parked:
                    where[where == -2] = -1; parked = False
                break
            # drop lk
            where[act[lk][0]] = -1
            R[:, lk:q - 1] = R[:, lk + 1:q]; R[:, q - 1] = 0
            u[lk:q - 1] = u[lk + 1:q]
            del act[lk]
            for k in range(lk, q - 1):
                where[act[k][0]] = k
            for j in range(lk, q - 1):
                a, b = R[j, j], R[j + 1, j]
                if b != 0:
                    h = np.hypot(a, b); cs, sn = a / h, b / h
                    G = np.array([[cs, sn], [-sn, cs]])
                    R[j:j + 2, j:q - 1] = G @ R[j:j + 2, j:q - 1]
                    N[:, j:j + 2] = N[:, j:j + 2] @ G.T
    

def solve_stack(desc, L, blob, active=None, tol_feas=1e-12, tol_rank=1e-11, levels=False):
    n, cin = desc.n, desc.c_in
    NR0 = n + cin
    eps = desc.eps
    lo = np.full(n, -INF); hi = np.full(n, INF)
    for s in range(desc.n_bound_sets):
        lo = np.maximum(lo, blob[L.l[s]:L.l[s] + n]); hi = np.minimum(hi, blob[L.u[s]:L.u[s] + n])
    Cin = blob[L.Cin:L.Cin + cin * n].reshape(cin, n) if cin else np.zeros((0, n))
    ylo0 = np.concatenate([clip(lo), clip(blob[L.lA:L.lA + cin]) if cin else []])
    yhi0 = np.concatenate([clip(hi), clip(blob[L.uA:L.uA + cin]) if cin else []])
    ZB = np.vstack([np.eye(n), Cin])
    z0, nz = 0, n
    yprev = np.zeros(NR0)
    status, iters = 0, 0
    xl = np.zeros((desc.n_levels, n))
    if desc.c_eq > 0:
        ce = desc.c_eq
        M = blob[L.Ceq:L.Ceq + ce * n].reshape(ce, n).copy()
        deq = blob[L.deq:L.deq + ce]
        rank, piv = sweep(ZB, 0, n, M, tol_rank)
        t = np.zeros(n); npiv = 0
        for r in range(ce):
            pc = piv[r]
            upto = pc if pc >= 0 else npiv
            s = deq[r] - M[r, :upto] @ t[:upto]
            if pc >= 0:
                t[pc] = s / M[r, pc]; npiv += 1
            elif abs(s) > 1e-9 * max(1, abs(deq[r])):
                status = 2
        yprev = ZB[:, :rank] @ t[:rank]
        z0, nz = rank, n - rank
    act_levels = [k for k in range(desc.n_levels) if active is None or active[k]]
    last = act_levels[-1]
    for k in act_levels:
        if status:
            break
        kind, wk = desc.task_kind[k], desc.weight_kind[k]
        m = desc.m[k]; cl = desc.c_loc[k]
        if nz == 0:
            xl[k] = yprev[:n]; continue
        path = "trivial" if (desc.m_reg == 0 and kind == 1 and wk == 0) else \
               "lq" if (desc.m_reg == 0 and kind == 0 and wk != 2) else "dense"
        ylo, yhi, yp = ylo0.copy(), yhi0.copy(), yprev.copy()
        if cl:
            Cl = blob[L.Cl[k]:L.Cl[k] + cl * n].reshape(cl, n)
            ylo = np.concatenate([ylo, clip(blob[L.lCl[k]:L.lCl[k] + cl])])
            yhi = np.concatenate([yhi, clip(blob[L.uCl[k]:L.uCl[k] + cl])])
            yp = np.concatenate([yp, Cl @ yprev[:n]])
        M = None
        rank_lq = 0
        if path == "trivial":
            Z = ZB[:n, z0:z0 + nz]
            b = blob[L.b[k]:L.b[k] + n]
            gz = Z.T @ ((1 + eps) * yprev[:n] - b + (blob[L.c[k]:L.c[k] + n] if desc.has_c[k] else 0))
            Lc = np.sqrt(1 + eps) * np.eye(nz)
        elif path == "lq":
            A = blob[L.A[k]:L.A[k] + m * n].reshape(m, n); b = blob[L.b[k]:L.b[k] + m]
            wd = blob[L.w[k]:L.w[k] + m] if wk == 1 else np.ones(m)
            M = A @ ZB[:n, z0:z0 + nz]; res = b - A @ yprev[:n]
            rank_lq, piv = sweep(ZB, z0, nz, M, tol_rank)
            for r in range(m):
                if piv[r] < 0:
                    pass
            Z = ZB[:n, z0:z0 + nz]
            Lm = M[:, :rank_lq]
            G = Lm.T @ (wd[:, None] * Lm) + eps * np.eye(rank_lq)
            gz = eps * (Z.T @ yprev[:n])
            if desc.has_c[k]:
                gz += Z.T @ blob[L.c[k]:L.c[k] + n]
            gz[:rank_lq] -= Lm.T @ (wd * res)
            Lc = np.sqrt(eps) * np.eye(nz)
            if rank_lq:
                Lc[:rank_lq, :rank_lq] = np.linalg.cholesky(G)
        else:
            Z = ZB[:n, z0:z0 + nz]
            Hz = eps * np.eye(nz)
            gz = eps * (Z.T @ yprev[:n])
            if desc.has_c[k]:
                gz += Z.T @ blob[L.c[k]:L.c[k] + n]
            if kind == 2:
                H = blob[L.A[k]:L.A[k] + n * n].reshape(n, n); g = blob[L.b[k]:L.b[k] + n]
                Hz += Z.T @ H @ Z; gz += Z.T @ (g + H @ yprev[:n])
            else:
                A = blob[L.A[k]:L.A[k] + m * n].reshape(m, n) if kind == 0 else np.eye(n)
                b = blob[L.b[k]:L.b[k] + m]
                if wk == 0: Wm = np.eye(m)
                elif wk == 1: Wm = np.diag(blob[L.w[k]:L.w[k] + m])
                else: Wm = blob[L.w[k]:L.w[k] + m * m].reshape(m, m)
                M = A @ Z; res = b - A @ yprev[:n]
                Hz += M.T @ Wm @ M; gz -= M.T @ (Wm @ res)
            if desc.m_reg:
                mr = desc.m_reg
                Ar = blob[L.Ar:L.Ar + mr * n].reshape(mr, n); br = blob[L.br:L.br + mr]
                wr = blob[L.wr:L.wr + mr] if desc.reg_weight_kind == 1 else np.ones(mr)
                Mr = Ar @ Z; rr = br - Ar @ yprev[:n]
                Hz += Mr.T @ (wr[:, None] * Mr); gz -= Mr.T @ (wr * rr)
            try:
                Lc = np.linalg.cholesky(Hz)
            except np.linalg.LinAlgError:
                status = 3 | (k << 8); break
        E = ZB[:, z0:z0 + nz].copy()
        if cl:
            E = np.vstack([E, Cl @ Z])
        p = np.linalg.solve(Lc, gz)
        N = np.linalg.solve(Lc, E.T).T     # E L^-T
        y = yp - N @ p
        st, it, act, where = gi(N, y, ylo, yhi, nz, tol_feas, 10 * (len(y) + n) + 20)
        iters += it
        if st:
            status = st | (k << 8); break
        xk = y[:n].copy()
        for (row, sg) in act:
            if row < n:
                xk[row] = ylo[row] if sg > 0 else yhi[row]
        yprev = np.concatenate([xk, y[n:NR0]])
        xl[k] = xk
        if path == "lq":
            z0, nz = z0 + rank_lq, nz - rank_lq
        elif k != last:
            if kind == 1:
                z0, nz = z0 + nz, 0
            elif kind == 0:
                rank, _ = sweep(ZB, z0, nz, M.copy(), tol_rank)
                z0, nz = z0 + rank, nz - rank
    return (yprev[:n], xl, status, iters) if levels else (yprev[:n], status, iters)

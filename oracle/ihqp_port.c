/*
 * ihqp_port.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see ihqp_port.h for who may load it and for the parity-pinning statement).
 *
 * What is restated, and from where (paths relative to the OpenSoT tree):
 *   port_cost_function   iHQP::computeCostFunction              src/solvers/iHQP.cpp:129-162
 *   bounds intersection  constraints::Aggregated::generateAll   src/constraints/Aggregated.cpp:141-148
 *   equality -> lA==uA   constraints::Aggregated::generateAll   src/constraints/Aggregated.cpp:159-172
 *   port_solve_stack     iHQP::solve level loop, optimality rows A_j x = A_j x_j*, piling
 *                                                                 src/solvers/iHQP.cpp:164-170,263-358
 *   eps on the diagonal, +-1e20 clipping   QPOasesBackEnd::solve  src/solvers/QPOasesBackEnd.cpp:251-255,339-356
 *   port_qp_solve        the job of BackEnd::solve. The reference's production back-end
 *        is qpOASES 3.1 (vendored, 26 kLoC, online active set); restating it is out of
 *        reach, so the single-QP solve restates the published Goldfarb-Idnani dual active
 *        set method that the reference's second back-end uses
 *        (external/eiQuadProg-ext/include/eiquadprog.hpp:139 solve_quadprog, steps 1/2a/2b/2c
 *        at :287-470), which the reference's own tests hold equal to qpOASES to 1e-12
 *        (tests/solvers/TesteiQuadProg.cpp:109-130). Differences from eiquadprog, all on
 *        purpose: two-sided rows and bounds are handled natively (at most one side of a
 *        row can be active) instead of being expanded to 2(n+m) one-sided rows
 *        (eiQuadProgBackEnd.cpp:42-68); a linearly dependent but consistent equality row
 *        is skipped instead of being reported as infeasible (the failure mode measured in
 *        SURVEY.md section 7); optimality is declared on a per-row feasibility tolerance
 *        instead of the cond(G)-scaled sum of infeasibilities (eiquadprog.hpp:291).
 * The strictly convex QP has a unique minimiser, so any correct method yields the same x.
 */
#include "ihqp_port.h"
#include "opensot_b200_synth.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../opensot_b200/csrc/osb_layout.c"

#define DEP_TOL   1e-22   /* |d2|^2 <= DEP_TOL |d|^2  ->  normal depends on the working set */
#define FEAS_TOL  1e-12
#define DEGEN_TOL 1e-8    /* largest violation attributed to rounding at a degenerate vertex  */
#define EQ_TOL    1e-9    /* consistency of a dependent equality row                        */

void port_cost_function(int m, int n, const double* A, const double* b, int wkind,
                        const double* W, const double* c, double* H, double* g)
{
    /* WA = W*A, Wb = W*b (Task::getWA/getWb, include/OpenSoT/Task.h:273-300) */
    double* WA = (double*)malloc(sizeof(double) * (size_t)(m * n + m + 1));
    double* Wb = WA + (size_t)m * n;
    for (int i = 0; i < m; ++i) {
        if (wkind == OSB_W_DENSE) {
            for (int j = 0; j < n; ++j) {
                double s = 0;
                for (int k = 0; k < m; ++k) s += W[i * m + k] * A[k * n + j];
                WA[i * n + j] = s;
            }
            double s = 0;
            for (int k = 0; k < m; ++k) s += W[i * m + k] * b[k];
            Wb[i] = s;
        } else {
            const double w = (wkind == OSB_W_DIAG) ? W[i] : 1.0;
            for (int j = 0; j < n; ++j) WA[i * n + j] = w * A[i * n + j];
            Wb[i] = w * b[i];
        }
    }
    for (int i = 0; i < n; ++i) {
        for (int j = i; j < n; ++j) { /* upper triangle then mirrored, iHQP.cpp:152-153 */
            double s = 0;
            for (int k = 0; k < m; ++k) s += A[k * n + i] * WA[k * n + j];
            H[i * n + j] = s;
            H[j * n + i] = s;
        }
        double s = 0;
        for (int k = 0; k < m; ++k) s += A[k * n + i] * Wb[k];
        g[i] = -s + (c ? c[i] : 0.0);
    }
    free(WA);
}

/* ------------------------------ Goldfarb-Idnani -------------------------------------- */
typedef struct gi_ws {
    int n, nb, nc, nt;
    const double* C;
    double *lo, *hi;          /* nt limits, clipped to +-1e20 */
    double *J, *R;            /* n x n row-major */
    double *d, *z, *r, *u, *np;
    int *act, *sgn, *where;   /* active list, its sign, position of constraint in list or -1 */
    int q, me;
} gi_ws;

static double con_value(const gi_ws* w, int i, const double* x)
{
    if (i < w->nb) return x[i];
    const double* row = w->C + (size_t)(i - w->nb) * w->n;
    double s = 0;
    for (int j = 0; j < w->n; ++j) s += row[j] * x[j];
    return s;
}
static void con_normal(const gi_ws* w, int i, int sign, double* np)
{
    if (i < w->nb) {
        memset(np, 0, sizeof(double) * w->n);
        np[i] = sign;
    } else {
        const double* row = w->C + (size_t)(i - w->nb) * w->n;
        for (int j = 0; j < w->n; ++j) np[j] = sign * row[j];
    }
}
/* d = J' np ; z = J2 d2 ; r = R^-1 d1  (eiquadprog.hpp:113-126) */
static void directions(gi_ws* w, double* dd, double* zz)
{
    const int n = w->n, q = w->q;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += w->J[i * n + j] * w->np[i];
        w->d[j] = s;
    }
    double a = 0, b = 0;
    for (int j = 0; j < n; ++j) { a += w->d[j] * w->d[j]; if (j >= q) b += w->d[j] * w->d[j]; }
    *dd = a; *zz = b;
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = q; j < n; ++j) s += w->J[i * n + j] * w->d[j];
        w->z[i] = s;
    }
    for (int i = q - 1; i >= 0; --i) {
        double s = w->d[i];
        for (int j = i + 1; j < q; ++j) s -= w->R[i * n + j] * w->r[j];
        w->r[i] = s / w->R[i * n + i];
    }
}
/* append np to the working set: rotate d2 onto e_q with Givens on J's columns
 * (eiquadprog.hpp add_constraint) */
static void add_constraint(gi_ws* w, int idx, int sign, double mult)
{
    const int n = w->n, q = w->q;
    for (int j = n - 1; j > q; --j) {
        double a = w->d[j - 1], b = w->d[j];
        if (b == 0.0) continue;
        double h = hypot(a, b), cc = a / h, ss = b / h;
        w->d[j - 1] = h; w->d[j] = 0.0;
        for (int i = 0; i < n; ++i) {
            double ja = w->J[i * n + j - 1], jb = w->J[i * n + j];
            w->J[i * n + j - 1] = cc * ja + ss * jb;
            w->J[i * n + j] = -ss * ja + cc * jb;
        }
    }
    for (int i = 0; i <= q; ++i) w->R[i * n + q] = w->d[i];
    w->act[q] = idx; w->sgn[q] = sign; w->u[q] = mult; w->where[idx] = q;
    w->q = q + 1;
}
/* remove position k from the working set (eiquadprog.hpp delete_constraint) */
static void drop_constraint(gi_ws* w, int k)
{
    const int n = w->n, q = w->q;
    w->where[w->act[k]] = -1;
    for (int j = k; j < q - 1; ++j) {
        for (int i = 0; i <= j + 1; ++i) w->R[i * n + j] = w->R[i * n + j + 1];
        w->act[j] = w->act[j + 1]; w->sgn[j] = w->sgn[j + 1]; w->u[j] = w->u[j + 1];
        w->where[w->act[j]] = j;
    }
    for (int j = k; j < q - 1; ++j) { /* R is upper Hessenberg from column k on */
        double a = w->R[j * n + j], b = w->R[(j + 1) * n + j];
        if (b == 0.0) continue;
        double h = hypot(a, b), cc = a / h, ss = b / h;
        for (int c = j; c < q - 1; ++c) {
            double ra = w->R[j * n + c], rb = w->R[(j + 1) * n + c];
            w->R[j * n + c] = cc * ra + ss * rb;
            w->R[(j + 1) * n + c] = -ss * ra + cc * rb;
        }
        for (int i = 0; i < n; ++i) {
            double ja = w->J[i * n + j], jb = w->J[i * n + j + 1];
            w->J[i * n + j] = cc * ja + ss * jb;
            w->J[i * n + j + 1] = -ss * ja + cc * jb;
        }
    }
    w->q = q - 1;
}

int port_qp_solve(int n, const double* H, const double* g, int nc, const double* C,
                  const double* lC, const double* uC, const double* l, const double* u,
                  double eps, double* x, int32_t* iters)
{
    gi_ws w;
    memset(&w, 0, sizeof(w));
    w.n = n; w.nb = (l && u) ? n : 0; w.nc = nc; w.nt = w.nb + nc; w.C = C;
    const int nt = w.nt;
    double* mem = (double*)calloc((size_t)(2 * n * n + 2 * nt + 6 * n + 2 * nt + n * n + 8), sizeof(double));
    int* imem = (int*)calloc((size_t)(2 * n + nt + 8), sizeof(int));
    double* p = mem;
    w.J = p; p += n * n; w.R = p; p += n * n;
    double* Lc = p; p += n * n;
    w.lo = p; p += nt; w.hi = p; p += nt;
    w.d = p; p += n; w.z = p; p += n; w.r = p; p += n; w.u = p; p += n + 1; w.np = p; p += n;
    w.act = imem; w.sgn = imem + n + 1; w.where = imem + 2 * n + 2;
    int status = OSB_OK, it = 0;

    for (int i = 0; i < nt; ++i) { /* checkINFTY, QPOasesBackEnd.cpp:339-356 */
        double lo = (i < w.nb) ? l[i] : lC[i - w.nb], hi = (i < w.nb) ? u[i] : uC[i - w.nb];
        w.lo[i] = lo < -OSB_INFTY ? -OSB_INFTY : lo;
        w.hi[i] = hi > OSB_INFTY ? OSB_INFTY : hi;
        w.where[i] = -1;
    }
    /* Cholesky of H + eps I (QPOasesBackEnd.cpp:253-255 adds eps to the diagonal) */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = H[i * n + j] + (i == j ? eps : 0.0);
            for (int k = 0; k < j; ++k) s -= Lc[i * n + k] * Lc[j * n + k];
            if (i == j) {
                if (!(s > 0.0)) { status = OSB_NOT_SPD; goto done; }
                Lc[i * n + i] = sqrt(s);
            } else Lc[i * n + j] = s / Lc[j * n + j];
        }
    /* J = L^-T (upper triangular): solve L' J = I column by column */
    for (int c = 0; c < n; ++c)
        for (int i = n - 1; i >= 0; --i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int k = i + 1; k < n; ++k) s -= Lc[k * n + i] * w.J[k * n + c];
            w.J[i * n + c] = s / Lc[i * n + i];
        }
    /* x = -H^-1 g by forward/back substitution with L (backward stable; going through the
     * explicit inverse J would amplify rounding by cond(L) for eps-regularised rank-deficient
     * Hessians, e.g. the KAT tests/solvers/TestQPOases.cpp:274-303 with eps = 2.2e-13) */
    for (int i = 0; i < n; ++i) {
        double s = -g[i];
        for (int k = 0; k < i; ++k) s -= Lc[i * n + k] * w.d[k];
        w.d[i] = s / Lc[i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = w.d[i];
        for (int k = i + 1; k < n; ++k) s -= Lc[k * n + i] * x[k];
        x[i] = s / Lc[i * n + i];
    }
    const int max_iter = 50 * (n + nt) + 100;
    int parked = 0;
    /* equality rows first (eiquadprog.hpp:232-275) */
    for (int i = 0; i < nt; ++i) {
        if (!(w.lo[i] == w.hi[i])) continue;
        double dd, zz;
        con_normal(&w, i, +1, w.np);
        directions(&w, &dd, &zz);
        double s = con_value(&w, i, x) - w.lo[i];
        if (zz <= DEP_TOL * dd || w.q == n) {
            if (fabs(s) <= EQ_TOL * fmax(1.0, fabs(w.lo[i]))) continue; /* redundant row */
            status = OSB_INFEASIBLE; goto done;
        }
        double t2 = -s / zz;
        for (int j = 0; j < n; ++j) x[j] += t2 * w.z[j];
        for (int j = 0; j < w.q; ++j) w.u[j] -= t2 * w.r[j];
        add_constraint(&w, i, +1, t2);
    }
    w.me = w.q;
    for (;;) {
        /* step 1: most violated inactive row (eiquadprog.hpp:279-320) */
        int ip = -1, sign = 0; double worst = 0.0;
        for (int i = 0; i < nt; ++i) {
            if (w.where[i] != -1 || w.lo[i] == w.hi[i]) continue;
            double v = con_value(&w, i, x);
            if (w.lo[i] > -OSB_INFTY) {
                double viol = (w.lo[i] - v) / fmax(1.0, fabs(w.lo[i]));
                if (viol > FEAS_TOL && viol > worst) { worst = viol; ip = i; sign = +1; }
            }
            if (w.hi[i] < OSB_INFTY) {
                double viol = (v - w.hi[i]) / fmax(1.0, fabs(w.hi[i]));
                if (viol > FEAS_TOL && viol > worst) { worst = viol; ip = i; sign = -1; }
            }
        }
        if (ip < 0) break;
        double up = 0.0;
        con_normal(&w, ip, sign, w.np);
        for (;;) { /* step 2a-2c */
            if (++it > max_iter) { status = OSB_MAX_ITER; goto done; }
            double dd, zz;
            directions(&w, &dd, &zz);
            double lim = (sign > 0) ? w.lo[ip] : w.hi[ip];
            double s = sign * (con_value(&w, ip, x) - lim); /* < 0 while violated */
            double t1 = INFINITY; int lk = -1;
            for (int k = w.me; k < w.q; ++k)
                if (w.r[k] > 0.0 && w.u[k] / w.r[k] < t1) { t1 = w.u[k] / w.r[k]; lk = k; }
            double t2 = (zz > DEP_TOL * dd && w.q < n) ? -s / zz : INFINITY;
            double t = fmin(t1, t2);
            if (!(t < INFINITY)) {
                /* degenerate vertex: a row violated by rounding only is parked until the
                 * iterate moves again; anything larger is a real inconsistency */
                if (-s <= DEGEN_TOL * fmax(1.0, fabs(lim))) { w.where[ip] = -2; parked = 1; break; }
                status = OSB_INFEASIBLE; goto done;
            }
            if (!(t2 < INFINITY)) { /* dual step only */
                for (int j = 0; j < w.q; ++j) w.u[j] -= t * w.r[j];
                up += t;
                drop_constraint(&w, lk);
                continue;
            }
            for (int j = 0; j < n; ++j) x[j] += t * w.z[j];
            for (int j = 0; j < w.q; ++j) w.u[j] -= t * w.r[j];
            up += t;
            if (t2 <= t1) {
                add_constraint(&w, ip, sign, up);
                if (parked) { for (int i = 0; i < nt; ++i) if (w.where[i] == -2) w.where[i] = -1; parked = 0; }
                break;
            }
            drop_constraint(&w, lk);
        }
    }
done:
    if (iters) *iters = it;
    free(mem); free(imem);
    return status;
}

/* -------------------------------- iHQP cascade --------------------------------------- */
int port_solve_stack(const osb_stack_desc* d, const double* blob, const int32_t* active,
                     double* x, double* x_levels, int32_t* iters)
{
    osb_layout L;
    if (osb_blob_layout(d, &L)) return OSB_NAN_INPUT;
    const int n = d->n;
    int max_rows = d->c_eq + d->c_in, sum_m = 0, max_cloc = 0;
    for (int k = 0; k < d->n_levels; ++k) {
        sum_m += (d->task_kind[k] == OSB_TASK_HESSIAN) ? 0 : d->m[k];
        if (d->c_loc[k] > max_cloc) max_cloc = d->c_loc[k];
    }
    max_rows += sum_m + max_cloc;
    double* mem = (double*)calloc((size_t)(3 * n * n + 6 * n + (size_t)max_rows * (n + 2) + 8), sizeof(double));
    double *H = mem, *Hr = H + n * n, *Aid = Hr + n * n, *g = Aid + n * n, *gr = g + n,
           *lb = gr + n, *ub = lb + n, *xk = ub + n, *C = xk + 2 * n,
           *lC = C + (size_t)max_rows * n, *uC = lC + max_rows;
    int total_it = 0, status = OSB_OK;
    for (int i = 0; i < n; ++i) Aid[i * n + i] = 1.0;
    for (int i = 0; i < n; ++i) x[i] = 0.0;

    /* bounds: u = min u_s, l = max l_s (constraints/Aggregated.cpp:141-148) */
    const int has_bounds = d->n_bound_sets > 0;
    for (int i = 0; i < n; ++i) { lb[i] = -INFINITY; ub[i] = INFINITY; }
    for (int s = 0; s < d->n_bound_sets; ++s)
        for (int i = 0; i < n; ++i) {
            lb[i] = fmax(lb[i], blob[L.l[s] + i]);
            ub[i] = fmin(ub[i], blob[L.u[s] + i]);
        }
    if (d->m_reg > 0) /* user regularisation, added to every level (iHQP.cpp:265-278) */
        port_cost_function(d->m_reg, n, blob + L.Ar, blob + L.br,
                           d->reg_weight_kind, d->reg_weight_kind == OSB_W_DIAG ? blob + L.wr : 0,
                           0, Hr, gr);

    /* rows that every level sees: global inequalities then equalities as lA == uA */
    int base_rows = 0;
    for (int r = 0; r < d->c_in; ++r, ++base_rows) {
        memcpy(C + (size_t)base_rows * n, blob + L.Cin + (size_t)r * n, sizeof(double) * n);
        lC[base_rows] = blob[L.lA + r]; uC[base_rows] = blob[L.uA + r];
    }
    for (int r = 0; r < d->c_eq; ++r, ++base_rows) {
        memcpy(C + (size_t)base_rows * n, blob + L.Ceq + (size_t)r * n, sizeof(double) * n);
        lC[base_rows] = uC[base_rows] = blob[L.deq + r];
    }
    int opt_rows = base_rows; /* optimality rows of earlier levels pile up after the base */
    for (int k = 0; k < d->n_levels; ++k) {
        if (active && !active[k]) continue; /* iHQP.cpp:270, 351-354 */
        const int kind = d->task_kind[k], m = d->m[k];
        const double* A = (kind == OSB_TASK_IDENTITY) ? Aid : blob + L.A[k];
        if (kind == OSB_TASK_HESSIAN) {
            memcpy(H, blob + L.A[k], sizeof(double) * n * n);
            memcpy(g, blob + L.b[k], sizeof(double) * n);
        } else {
            port_cost_function(m, n, A, blob + L.b[k], d->weight_kind[k],
                               d->weight_kind[k] == OSB_W_IDENTITY ? 0 : blob + L.w[k],
                               d->has_c[k] ? blob + L.c[k] : 0, H, g);
        }
        if (d->m_reg > 0) {
            for (int i = 0; i < n * n; ++i) H[i] += Hr[i];
            for (int i = 0; i < n; ++i) g[i] += gr[i];
        }
        int rows = opt_rows; /* task-local rows go last for this level only */
        for (int r = 0; r < d->c_loc[k]; ++r, ++rows) {
            memcpy(C + (size_t)rows * n, blob + L.Cl[k] + (size_t)r * n, sizeof(double) * n);
            lC[rows] = blob[L.lCl[k] + r]; uC[rows] = blob[L.uCl[k] + r];
        }
        int it = 0;
        int st = port_qp_solve(n, H, g, rows, C, lC, uC, has_bounds ? lb : 0, has_bounds ? ub : 0,
                               d->eps, xk, &it);
        total_it += it;
        if (st != OSB_OK) { status = st | (k << 8); break; }
        memcpy(x, xk, sizeof(double) * n);
        if (x_levels) memcpy(x_levels + (size_t)k * n, xk, sizeof(double) * n);
        if (kind != OSB_TASK_HESSIAN) { /* optimality rows A_k x = A_k x_k* (iHQP.cpp:164-170) */
            for (int r = 0; r < m; ++r, ++opt_rows) {
                double s = 0;
                for (int j = 0; j < n; ++j) { C[(size_t)opt_rows * n + j] = A[r * n + j]; s += A[r * n + j] * xk[j]; }
                lC[opt_rows] = uC[opt_rows] = s;
            }
        }
    }
    if (iters) *iters = total_it;
    free(mem);
    return status;
}

int port_solve_batch(const osb_stack_desc* d, const double* blobs, int32_t batch,
                     const int32_t* active, double* x, double* x_levels, int32_t* status,
                     int32_t* iters)
{
    osb_layout L;
    if (osb_blob_layout(d, &L)) return OSB_E_INVALID;
    for (int i = 0; i < batch; ++i) {
        int32_t it = 0;
        int st = port_solve_stack(d, blobs + (size_t)i * L.stride, active, x + (size_t)i * d->n,
                                  x_levels ? x_levels + (size_t)i * d->n_levels * d->n : 0, &it);
        if (status) status[i] = st;
        if (iters) iters[i] = it;
    }
    return OSB_SUCCESS;
}

int port_synth_shape(const char* name, double eps_factor, osb_stack_desc* d)
{
    return synth_shape(name, eps_factor, d);
}
int port_blob_layout(const osb_stack_desc* d, osb_layout* L) { return osb_blob_layout(d, L); }
int port_synth_fill(const osb_stack_desc* d, uint64_t seed, uint64_t first, int32_t count,
                    double tight, double* blobs)
{
    osb_layout L;
    if (osb_blob_layout(d, &L)) return OSB_E_INVALID;
    for (int i = 0; i < count; ++i)
        synth_fill(d, &L, seed, first + (uint64_t)i, tight, blobs + (size_t)i * L.stride);
    return OSB_SUCCESS;
}

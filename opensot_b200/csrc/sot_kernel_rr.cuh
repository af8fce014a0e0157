/*
 * sot_kernel_rr.cuh -- the hot path with REGISTER-RESIDENT ROWS: one warp solves one stack,
 * and the matrix a phase transforms lives in the registers of the lanes that own its rows.
 *
 * Same mathematics, same reference semantics and same status words as sot_kernel.cuh (see
 * the header of that file and DESIGN.md section 3: null-space cascade instead of optimality
 * equalities, LQ-first level factorisation, Goldfarb-Idnani dual active set on the box
 * ylo <= y = y_prev + N z <= yhi). Replaces, per instance and per control cycle
 * (SURVEY.md section 8a):
 *   a-2  bounds intersection / +-1e20 clipping   src/constraints/Aggregated.cpp:141-148,
 *                                                src/solvers/QPOasesBackEnd.cpp:339-356
 *   a-5  H = A'WA + eps I, g = -A'Wb + c         src/solvers/iHQP.cpp:129-162,274-278
 *                                                src/solvers/QPOasesBackEnd.cpp:253-255
 *   a-6  optimality rows A_j x = A_j x_j*        src/solvers/iHQP.cpp:164-170,285-332
 *   a-8  the dense QP of every level             src/solvers/QPOasesBackEnd.cpp:248-307
 *   a-9  the level loop / active levels          src/solvers/iHQP.cpp:263-358
 *
 * What is different from sot_kernel.cuh is where the data lives and how the loops look:
 *  - the generic kernel walks shared-memory rows with rolled loops: ~170 k cycles per stack,
 *    bound by exposed shared-memory latency (8 resident warps per SM, measured);
 *  - here a lane keeps its row (of the task block M during the LQ sweep, of the basis ZB
 *    while it takes the reflectors, of N = (E Z) L^-T during the active-set loop) in
 *    registers as NP pairs, every inner loop is a fully unrolled walk over those NP pairs
 *    against a broadcast vector (one LDS.128 per pair, no address arithmetic, no loop
 *    branches, independent FMAs), and matrices that were only staging copies (T) are gone
 *    from shared memory, which raises the number of resident warps.
 *  - column indices are ABSOLUTE (0..LD): the live basis is columns [z0, n) of ZB; vectors
 *    are zero outside the live range, so no walk needs a mask or a peel.
 *  - the only dynamic column accesses (the pivot column of a Householder step) go through a
 *    switch over the NP register pairs (rr_get / rr_set).
 * Kernel instances: <Real, NP, RPL> = pairs per row, rows per lane; the host picks the
 * smallest that holds the stack (rr_plan_workspace), shapes outside fall to the generic kernel.
 */
#ifndef OSB_SOT_KERNEL_RR_CUH_
#define OSB_SOT_KERNEL_RR_CUH_

#ifndef OSB_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include "sot_params.h"
#include "sot_common.cuh"

#ifdef OSB_RR_PROFILE
#define RPROF_T(var) const long long var = clock64()
#else
#define RPROF_T(var)
#endif

namespace {

#define RR_CASES(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) \
    X(16) X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31)

/* element j of a register row (j is warp-uniform but dynamic): switch over the pairs */
template <typename Real, int NP>
__device__ __forceinline__ Real rr_get(const typename Tr<Real>::V2 (&t)[NP], int j)
{
    typename Tr<Real>::V2 v = t[0];
    switch (j >> 1) {
#define X(K) case K: if constexpr (K < NP) v = t[K]; break;
        RR_CASES(X)
#undef X
    default: break;
    }
    return (j & 1) ? v.y : v.x;
}
template <typename Real, int NP>
__device__ __forceinline__ void rr_set(typename Tr<Real>::V2 (&t)[NP], int j, Real val)
{
    const bool hi = (j & 1) != 0;
    switch (j >> 1) {
#define X(K) case K: if constexpr (K < NP) { if (hi) t[K].y = val; else t[K].x = val; } break;
        RR_CASES(X)
#undef X
    default: break;
    }
}
template <typename Real, int NP>
__device__ __forceinline__ void rr_load(typename Tr<Real>::V2 (&t)[NP], const Real* __restrict__ row)
{
    const typename Tr<Real>::V2* r2 = reinterpret_cast<const typename Tr<Real>::V2*>(row);
#pragma unroll
    for (int k = 0; k < NP; ++k) t[k] = r2[k];
}
template <typename Real, int NP>
__device__ __forceinline__ void rr_store(Real* __restrict__ row, const typename Tr<Real>::V2 (&t)[NP])
{
    typename Tr<Real>::V2* r2 = reinterpret_cast<typename Tr<Real>::V2*>(row);
#pragma unroll
    for (int k = 0; k < NP; ++k) r2[k] = t[k];
}
template <typename Real, int NP>
__device__ __forceinline__ void rr_zero(typename Tr<Real>::V2 (&t)[NP])
{
#pragma unroll
    for (int k = 0; k < NP; ++k) { t[k].x = Real(0.0); t[k].y = Real(0.0); }
}
/* static select of slot s out of a small per-slot register array */
template <typename T, int RPL>
__device__ __forceinline__ T rr_pick(const T (&a)[RPL], int s)
{
    T v = a[0];
#pragma unroll
    for (int k = 1; k < RPL; ++k) if (s == k) v = a[k];
    return v;
}

/* Everything one warp needs besides its register rows. */
template <typename Real>
struct RC {
    const sot_params* p;
    int lane;
    Real *blob, *ZB, *M, *V, *Rp;
    Real *yprev, *d, *dm, *pv, *cscale, *cs, *sn, *u, *r, *rinv, *res, *beta;
    int32_t *act, *piv;
    int n, cin, NR0, NR;
    int z0;             /* live basis = columns [z0, n) */
};

/* dot of two shared-memory runs of dynamic length (short: the leading block of a level) */
template <typename Real>
__device__ __forceinline__ Real rr_dot_dyn(const Real* __restrict__ a, const Real* __restrict__ b, int cnt)
{
    Real s0 = 0, s1 = 0;
    int j = 0;
#pragma unroll 2
    for (; j + 1 < cnt; j += 2) { s0 = fma(a[j], b[j], s0); s1 = fma(a[j + 1], b[j + 1], s1); }
    if (j < cnt) s0 = fma(a[j], b[j], s0);
    return s0 + s1;
}

/*
 * dst rows (row r at dst + r*LD, absolute columns, zero outside the live basis)
 *   <-  A (m x n row-major, row stride lda) times the basis.  Lane per basis column.
 */
template <typename Real, int LD>
__device__ void rr_project(RC<Real>& c, const Real* __restrict__ A, int lda, int m, Real* dst)
{
    const int n = c.n;
#pragma unroll 1
    for (int j = c.lane; j < LD; j += 32) {
        const bool live = (j >= c.z0) && (j < n);
        const Real* zcol = c.ZB + j;
        int r = 0;
#pragma unroll 1
        for (; r + 3 < m; r += 4) {
            Real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            if (live) {
                const Real *a0 = A + r * lda, *a1 = a0 + lda, *a2 = a1 + lda, *a3 = a2 + lda;
#pragma unroll 4
                for (int i = 0; i < n; ++i) {
                    const Real zv = zcol[i * LD];
                    s0 = fma(a0[i], zv, s0); s1 = fma(a1[i], zv, s1);
                    s2 = fma(a2[i], zv, s2); s3 = fma(a3[i], zv, s3);
                }
            }
            dst[r * LD + j] = s0; dst[(r + 1) * LD + j] = s1;
            dst[(r + 2) * LD + j] = s2; dst[(r + 3) * LD + j] = s3;
        }
#pragma unroll 1
        for (; r < m; ++r) {
            Real s0 = 0, s1 = 0;
            if (live) {
                const Real* a0 = A + r * lda;
                int i = 0;
#pragma unroll 2
                for (; i + 1 < n; i += 2) { s0 = fma(a0[i], zcol[i * LD], s0); s1 = fma(a0[i + 1], zcol[(i + 1) * LD], s1); }
                if (i < n) s0 = fma(a0[i], zcol[i * LD], s0);
            }
            dst[r * LD + j] = s0 + s1;
        }
    }
}
/* dst rows <- A with the row stride changed from lda to LD and zero padding (identity basis) */
template <typename Real, int LD>
__device__ __forceinline__ void rr_repack(RC<Real>& c, const Real* __restrict__ A, int lda, int m, Real* dst)
{
#pragma unroll 1
    for (int r = 0; r < m; ++r)
#pragma unroll 1
        for (int j = c.lane; j < LD; j += 32) dst[r * LD + j] = (j < c.n) ? A[r * lda + j] : Real(0.0);
}
/* out[r] = rhs[r] - A[r,:] . x   (lane per row) */
template <typename Real>
__device__ __forceinline__ void rr_residual(RC<Real>& c, const Real* __restrict__ A, int lda, const Real* rhs, int m,
                                            const Real* x, Real* out)
{
#pragma unroll 1
    for (int r = c.lane; r < m; r += 32)
        out[r] = (rhs ? rhs[r] : Real(0.0)) - rr_dot_dyn<Real>(A + r * lda, x, c.n);
}

/*
 * Householder sweep: pushes the row space of the m rows in c.M (absolute columns, zero outside
 * the live basis) out of the basis. Afterwards c.M holds [L 0] (pivot column of row r in
 * c.piv[r], -1 for a linearly dependent row), the basis rows ZB have taken the reflectors,
 * and the live columns [z0 + rank, n) span the null space. Returns the rank.
 * Replaces adding the optimality equalities of a finished level to every later QP
 * (iHQP.cpp:164-170, 314-332); dependent rows (|residual| <= tol_rank |row|) are skipped, the
 * job of qpOASES' linear-independence test (QProblem.cpp:2847).
 *
 * Phase 1, LQ of M: lane r owns row r in registers. Step r: the owner publishes its row a
 * (zero before the pivot column jz: finished entries have been moved out) as V[zc] and
 * clears its registers; every lane walks V[zc] once for its own a.x and (redundantly) |a|^2,
 * forms rho, beta of H = I - beta u u', u = a - rho e_jz, and takes the reflector with a
 * second walk; its pivot-column entry is the finished L[row][zc]: stored to c.M and cleared.
 * Phase 2: every basis row (lane per row, registers) takes u_0 .. u_{rank-1} back to back.
 */
template <typename Real, int NP, int RPL>
__device__ int rr_sweep(RC<Real>& c, typename Tr<Real>::V2 (&t)[RPL][NP], int m)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP;
    const int lane = c.lane, z0 = c.z0, n = c.n;
    const Real tol2 = (Real)(c.p->tol_rank * c.p->tol_rank);
    Real rn2[RPL];
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        rn2[s] = Real(0.0);
        if (row < m) {
            rr_load<Real, NP>(t[s], c.M + row * LD);
            Real a0 = 0, a1 = 0;
#pragma unroll
            for (int k = 0; k < NP; ++k) { a0 = fma(t[s][k].x, t[s][k].x, a0); a1 = fma(t[s][k].y, t[s][k].y, a1); }
            rn2[s] = a0 + a1;
            V2 zero; zero.x = Real(0.0); zero.y = Real(0.0);
            V2* r2 = reinterpret_cast<V2*>(c.M + row * LD);
#pragma unroll
            for (int k = 0; k < NP; ++k) r2[k] = zero;
            c.piv[row] = -1;
        } else {
            rr_zero<Real, NP>(t[s]);
        }
    }
    __syncwarp();
    int zc = 0;
#pragma unroll 1
    for (int r = 0; r < m && z0 + zc < n; ++r) {
        const int jz = z0 + zc, ol = r & 31, os = r >> 5;
        Real* Vz = c.V + zc * LD;
        if (lane == ol) {
#pragma unroll
            for (int s = 0; s < RPL; ++s)
                if (s == os) { rr_store<Real, NP>(Vz, t[s]); rr_zero<Real, NP>(t[s]); }
        }
        __syncwarp();
        const V2* v2 = reinterpret_cast<const V2*>(Vz);
        Real dx[RPL], dy[RPL], n0 = 0, n1 = 0;
#pragma unroll
        for (int s = 0; s < RPL; ++s) { dx[s] = Real(0.0); dy[s] = Real(0.0); }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const V2 av = v2[k];
            n0 = fma(av.x, av.x, n0); n1 = fma(av.y, av.y, n1);
#pragma unroll
            for (int s = 0; s < RPL; ++s) { dx[s] = fma(av.x, t[s][k].x, dx[s]); dy[s] = fma(av.y, t[s][k].y, dy[s]); }
        }
        const Real nn = n0 + n1;
        const Real a_p = Vz[jz];
        const Real rn2o = __shfl_sync(FULL, rr_pick<Real, RPL>(rn2, os), ol);
        if (!(nn > tol2 * rn2o) || nn <= Real(0.0)) { /* dependent (or zero) row: dropped, its registers are clear */
            __syncwarp();
            continue;
        }
        const Real rs = rsqrt(nn), sq = nn * rs;
        const Real rho = (a_p > Real(0.0)) ? -sq : sq;
        const Real beta = Tr<Real>::rcp(fma(sq, fabs(a_p), nn));   /* 2 / |u|^2 */
        Real w[RPL], lv[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const Real xp = rr_get<Real, NP>(t[s], jz);
            w[s] = beta * ((dx[s] + dy[s]) - rho * xp);
            lv[s] = xp - w[s] * (a_p - rho);
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const V2 av = v2[k];
#pragma unroll
            for (int s = 0; s < RPL; ++s) { t[s][k].x = fma(-w[s], av.x, t[s][k].x); t[s][k].y = fma(-w[s], av.y, t[s][k].y); }
        }
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int row = lane + 32 * s;
            if (row > r && row < m) c.M[row * LD + jz] = lv[s];
            rr_set<Real, NP>(t[s], jz, Real(0.0));
        }
        __syncwarp();                       /* every lane has walked V[zc] twice */
        if (lane == 0) { c.M[r * LD + jz] = rho; c.piv[r] = zc; c.beta[zc] = beta; Vz[jz] = a_p - rho; }
        ++zc;
    }
    __syncwarp();
    const int rank = zc;
    /* phase 2: basis rows (x rows and Cin rows) take all reflectors */
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        if (row < c.NR0) rr_load<Real, NP>(t[s], c.ZB + row * LD);
        else rr_zero<Real, NP>(t[s]);
    }
#pragma unroll 1
    for (int k = 0; k < rank; ++k) {
        const V2* v2 = reinterpret_cast<const V2*>(c.V + k * LD);
        const Real beta = c.beta[k];
        Real dx[RPL], dy[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) { dx[s] = Real(0.0); dy[s] = Real(0.0); }
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) {
            const V2 av = v2[kk];
#pragma unroll
            for (int s = 0; s < RPL; ++s) { dx[s] = fma(av.x, t[s][kk].x, dx[s]); dy[s] = fma(av.y, t[s][kk].y, dy[s]); }
        }
        Real w[RPL];
#pragma unroll
        for (int s = 0; s < RPL; ++s) w[s] = beta * (dx[s] + dy[s]);
#pragma unroll
        for (int kk = 0; kk < NP; ++kk) {
            const V2 av = v2[kk];
#pragma unroll
            for (int s = 0; s < RPL; ++s) { t[s][kk].x = fma(-w[s], av.x, t[s][kk].x); t[s][kk].y = fma(-w[s], av.y, t[s][kk].y); }
        }
    }
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        if (row < c.NR0) rr_store<Real, NP>(c.ZB + row * LD, t[s]);
    }
    __syncwarp();
    return rank;
}

struct RRLevelOut { int status; int iters; };

/* offset of column j of the packed working-set factor: column j has j + 2 slots (rows 0..j
 * and the sub-diagonal entry that exists while a member is being removed) */
__device__ __forceinline__ int rr_col(int j) { return (j * (j + 3)) >> 1; }

/*
 * Goldfarb-Idnani dual active-set iteration on the box lo <= y <= hi, y = y0 + N z, one row of
 * N per lane and slot in registers (t), working-set factor R (q x q upper, packed by columns
 * in c.Rp). The single-QP job of BackEnd::solve (QPOasesBackEnd.cpp:248-307); method after
 * Goldfarb & Idnani 1983 as used by the reference's eiQuadProg back-end
 * (external/eiQuadProg-ext/include/eiquadprog.hpp:287-470). Working-set member k occupies
 * column z0 + k of N; free columns are [z0 + q, n).
 */
template <typename Real, int NP, int RPL>
__device__ RRLevelOut rr_active_set(RC<Real>& c, typename Tr<Real>::V2 (&t)[RPL][NP], Real (&yv)[RPL],
                                    const Real (&lo)[RPL], const Real (&hi)[RPL], const Real (&rq)[RPL],
                                    int (&wh)[RPL], int max_iter)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP;
    const int lane = c.lane, z0 = c.z0, n = c.n, nz = n - z0;
    const Real tol = (Real)c.p->tol_feas;
    RRLevelOut out{OSB_OK, 0};
    bool parked = false;
    int q = 0;
#pragma unroll
    for (int s = 0; s < RPL; ++s) wh[s] = -1;
#pragma unroll 1
    for (;;) {
        /* ---- pricing: most violated row that is neither in the working set nor parked ---- */
        Real best = Real(0.0); int bidx = -1;
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int row = lane + 32 * s;
            const Real v = yv[s], l = lo[s], h = hi[s];
            Real vl = (l - v) - tol * fmax(Real(1.0), fabs(l)); /* > 0: violated beyond tolerance */
            Real vh = (v - h) - tol * fmax(Real(1.0), fabs(h));
            vl = (l > -Real(OSB_INFTY)) ? vl : -Real(1.0);
            vh = (h < Real(OSB_INFTY)) ? vh : -Real(1.0);
            const bool upper = vh > vl;
            const Real cand = upper ? vh : vl;
            if (wh[s] == -1 && cand > best) {
                best = cand;
                bidx = row | (upper ? 0x10000 : 0) | ((l == h) ? 0x20000 : 0);
            }
        }
        bidx = warp_argmax_nonneg(best, bidx);
        if (!(best > Real(0.0))) break; /* primal feasible: optimal */
        const int ip = bidx & 0xffff, ipl = ip & 31, ips = ip >> 5;
        const bool up_side = (bidx & 0x10000) != 0;
        const Real sgn = up_side ? -Real(1.0) : Real(1.0); /* +1: lower side, -1: upper side */
        const Real lim = __shfl_sync(FULL, up_side ? rr_pick<Real, RPL>(hi, ips) : rr_pick<Real, RPL>(lo, ips), ipl);
        const Real dd = __shfl_sync(FULL, rr_pick<Real, RPL>(rq, ips), ipl);
        Real up = Real(0.0);
#pragma unroll 1
        for (;;) {
            if (++out.iters > max_iter) { out.status = OSB_MAX_ITER; return out; }
            const int jq = z0 + q;
            /* d = N' n_p: the owner publishes its row; lanes over columns sign it and mask it
             * to the free columns (dm), |d2|^2 on the side */
            if (lane == ipl) {
#pragma unroll
                for (int s = 0; s < RPL; ++s) if (s == ips) rr_store<Real, NP>(c.d, t[s]);
            }
            __syncwarp();
            Real zp = Real(0.0);
#pragma unroll
            for (int w0 = 0; w0 < LD; w0 += 32) {
                const int w = w0 + lane;
                if (w < LD) {
                    const Real dj = sgn * c.d[w];
                    const Real mj = (w >= jq) ? dj : Real(0.0);
                    c.d[w] = dj; c.dm[w] = mj;
                    zp = fma(mj, mj, zp);
                }
            }
            __syncwarp();
            const Real zz = warp_sum<Real>(zp);
            /* dy = N2 d2 for every row */
            Real dyv[RPL];
            {
                const V2* dv = reinterpret_cast<const V2*>(c.dm);
                Real ax[RPL], ay[RPL];
#pragma unroll
                for (int s = 0; s < RPL; ++s) { ax[s] = Real(0.0); ay[s] = Real(0.0); }
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    const V2 v = dv[k];
#pragma unroll
                    for (int s = 0; s < RPL; ++s) { ax[s] = fma(t[s][k].x, v.x, ax[s]); ay[s] = fma(t[s][k].y, v.y, ay[s]); }
                }
#pragma unroll
                for (int s = 0; s < RPL; ++s) dyv[s] = ax[s] + ay[s];
            }
            /* r = R^-1 d1 (column-oriented back substitution, one value per lane) */
            Real t1 = INFINITY; int lk = -1;
            if (q > 0) {
                Real r0 = (lane < q) ? c.d[z0 + lane] : Real(0.0);
                Real r1 = (lane + 32 < q) ? c.d[z0 + lane + 32] : Real(0.0);
                const Real ri0 = (lane < q) ? c.rinv[lane] : Real(0.0);
                const Real ri1 = (lane + 32 < q) ? c.rinv[lane + 32] : Real(0.0);
                Real col0 = (lane < q - 1) ? c.Rp[rr_col(q - 1) + lane] : Real(0.0);
#pragma unroll 1
                for (int j = q - 1; j >= 0; --j) {
                    const Real nxt = (j > 0 && lane < j - 1) ? c.Rp[rr_col(j - 1) + lane] : Real(0.0);
                    const Real mine = (j < 32) ? r0 * ri0 : r1 * ri1;
                    const Real rj = __shfl_sync(FULL, mine, j & 31);
                    if (lane == (j & 31)) { if (j < 32) r0 = rj; else r1 = rj; }
                    r0 = fma(-col0, rj, r0);
                    if (lane + 32 < j) r1 = fma(-c.Rp[rr_col(j) + lane + 32], rj, r1);
                    col0 = nxt;
                }
                if (lane < q) c.r[lane] = r0;
                if (lane + 32 < q) c.r[lane + 32] = r1;
                /* ratio test over droppable (non-equality) working-set members */
#pragma unroll 1
                for (int k = lane; k < q; k += 32) {
                    const Real rk = (k < 32) ? r0 : r1;
                    if (rk > Real(0.0) && !(c.act[k] & 0x20000)) {
                        const Real ratio = fmax(c.u[k] / rk, Real(0.0)); /* a multiplier is never negative */
                        if (ratio < t1) { t1 = ratio; lk = k; }
                    }
                }
                lk = warp_argmin_nonneg(t1, lk);
            }
            __syncwarp();
            const Real y_ip = __shfl_sync(FULL, rr_pick<Real, RPL>(yv, ips), ipl);
            const Real sv = sgn * (y_ip - lim); /* < 0 while violated */
            const bool indep = (zz > Tr<Real>::dep_tol * dd) && (q < nz);
            const Real t2 = indep ? -sv / zz : INFINITY;
            const Real tt = fmin(t1, t2);
            if (!(tt < INFINITY)) {
                /* neither a primal nor a dual step exists: a row violated by rounding only at a
                 * degenerate vertex is parked until the iterate moves; anything larger is a
                 * genuinely inconsistent constraint set (same rule as sot_kernel.cuh) */
                if (-sv <= Tr<Real>::degenerate_tol * fmax(Real(1.0), fabs(lim))) {
                    if (lane == ipl) {
#pragma unroll
                        for (int s = 0; s < RPL; ++s) if (s == ips) wh[s] = -2;
                    }
                    parked = true;
                    break;
                }
                out.status = OSB_INFEASIBLE; return out;
            }
            if (indep) { /* primal + dual step */
#pragma unroll
                for (int s = 0; s < RPL; ++s) yv[s] = fma(tt, dyv[s], yv[s]);
            }
#pragma unroll 1
            for (int k = lane; k < q; k += 32) c.u[k] = fma(-tt, c.r[k], c.u[k]);
            up += tt;
            if (indep && t2 <= t1) {
                /* ---- full step: append ip. Householder maps d2 onto rho e_q ---- */
                const Real dq = c.d[jq];
                const Real rs = rsqrt(zz), sq = zz * rs;         /* 1/|d2|, |d2| */
                const Real rho = (dq > Real(0.0)) ? -sq : sq;
                const Real rho_inv = (dq > Real(0.0)) ? -rs : rs;
                const Real beta = Tr<Real>::rcp(rho * (rho - dq));
                __syncwarp();                                       /* every lane has walked dm */
                if (lane == 0) c.dm[jq] = dq - rho;                 /* v = d2 - rho e1 */
                __syncwarp();
                {
                    const V2* dv = reinterpret_cast<const V2*>(c.dm);
                    Real w[RPL];
#pragma unroll
                    for (int s = 0; s < RPL; ++s) w[s] = beta * (dyv[s] - rho * rr_get<Real, NP>(t[s], jq)); /* (N2 v)_i */
#pragma unroll
                    for (int k = 0; k < NP; ++k) {
                        const V2 v = dv[k];
#pragma unroll
                        for (int s = 0; s < RPL; ++s) { t[s][k].x = fma(-w[s], v.x, t[s][k].x); t[s][k].y = fma(-w[s], v.y, t[s][k].y); }
                    }
                }
                const int cq = rr_col(q);
#pragma unroll 1
                for (int k = lane; k < q; k += 32) c.Rp[cq + k] = c.d[z0 + k];
                if (lane == 0) {
                    c.Rp[cq + q] = rho; c.rinv[q] = rho_inv;
                    c.act[q] = bidx; c.u[q] = up;
                }
                if (lane == ipl) {
#pragma unroll
                    for (int s = 0; s < RPL; ++s) if (s == ips) wh[s] = q;
                }
                ++q;
                if (parked) { /* the iterate moved: parked rows are priced again */
#pragma unroll
                    for (int s = 0; s < RPL; ++s) if (wh[s] == -2) wh[s] = -1;
                    parked = false;
                }
                __syncwarp();
                break;
            }
            __syncwarp();
            /* ---- partial (or pure dual) step: drop working-set member lk ---- */
            {
                const int k0 = lk;
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    if (wh[s] == k0) wh[s] = -1;
                    else if (wh[s] > k0) wh[s] -= 1;
                }
                /* delete column k0 of R: column c+1 becomes column c (lane per row) */
#pragma unroll 1
                for (int i = lane; i < q; i += 32) {
                    const int from = (i - 1 > k0) ? i - 1 : k0;
#pragma unroll 1
                    for (int cc = from; cc < q - 1; ++cc) c.Rp[rr_col(cc) + i] = c.Rp[rr_col(cc + 1) + i];
                }
                Real un = 0; int an = 0;
#pragma unroll 1
                for (int base = k0; base < q - 1; base += 32) {
                    const int k = base + lane;
                    if (k < q - 1) { un = c.u[k + 1]; an = c.act[k + 1]; }
                    __syncwarp();
                    if (k < q - 1) { c.u[k] = un; c.act[k] = an; }
                    __syncwarp();
                }
                __syncwarp();
                /* Givens rotations restore the triangle; the same rotations act on columns
                 * z0+j, z0+j+1 of N: their cosines/sines are kept for the register pass below */
#pragma unroll 1
                for (int j = k0; j < q - 1; ++j) {
                    const Real a = c.Rp[rr_col(j) + j], b = c.Rp[rr_col(j) + j + 1];
                    Real cs = Real(1.0), sn = Real(0.0);
                    if (b != Real(0.0)) {
                        const Real hinv = rsqrt(fma(a, a, b * b));
                        cs = a * hinv; sn = b * hinv;
                        __syncwarp();
#pragma unroll 1
                        for (int cc = j + lane; cc < q - 1; cc += 32) {
                            Real* pj = c.Rp + rr_col(cc) + j;
                            const Real ra = pj[0], rb = pj[1];
                            pj[0] = cs * ra + sn * rb;
                            pj[1] = cs * rb - sn * ra;
                        }
                        __syncwarp();
                    }
                    if (lane == 0) { c.rinv[j] = Tr<Real>::rcp(c.Rp[rr_col(j) + j]); c.cs[z0 + j] = cs; c.sn[z0 + j] = sn; }
                }
                __syncwarp();
                {
                    const int jlo = z0 + k0, jhi = z0 + q - 1;
#pragma unroll
                    for (int cj = 0; cj < LD - 1; ++cj) {
                        if (cj >= jlo && cj < jhi) {
                            const Real cs = c.cs[cj], sn = c.sn[cj];
#pragma unroll
                            for (int s = 0; s < RPL; ++s) {
                                Real& ea = (cj & 1) ? t[s][cj >> 1].y : t[s][cj >> 1].x;
                                Real& eb = ((cj + 1) & 1) ? t[s][(cj + 1) >> 1].y : t[s][(cj + 1) >> 1].x;
                                const Real ja = ea, jb = eb;
                                ea = cs * ja + sn * jb;
                                eb = cs * jb - sn * ja;
                            }
                        }
                    }
                }
                --q;
                __syncwarp();
            }
        }
    }
    return out;
}

template <typename Real, int NP, int RPL>
__device__ __forceinline__ void rr_solve_impl(const sot_params& p)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP;
    OSB_SMEM_RAW(ws_raw);
    Real* ws = reinterpret_cast<Real*>(ws_raw);
    const int lane = threadIdx.x;
    const int inst = blockIdx.x;
    if (inst >= p.batch) return;
    const osb_stack_desc& D = p.d;
    const osb_layout& L = p.L;
    const rr_workspace& W = p.rr;

    RC<Real> c;
    c.p = &p; c.lane = lane;
    c.blob = ws + W.off_blob; c.ZB = ws + W.off_ZB; c.M = ws + W.off_M; c.V = ws + W.off_V; c.Rp = ws + W.off_R;
    c.yprev = ws + W.off_yprev; c.d = ws + W.off_d; c.dm = ws + W.off_dm; c.pv = ws + W.off_p;
    c.cscale = ws + W.off_cscale; c.cs = ws + W.off_cs; c.sn = ws + W.off_sn;
    c.u = ws + W.off_u; c.r = ws + W.off_r; c.rinv = ws + W.off_rinv; c.res = ws + W.off_res; c.beta = ws + W.off_beta;
    c.act = (int32_t*)(ws + W.off_act); c.piv = (int32_t*)(ws + W.off_piv);
    const int stride_e = (sizeof(Real) == 8) ? L.stride : ((L.stride + 3) & ~3);
    Real* const out_x = static_cast<Real*>(p.x);
    Real* const out_xl = static_cast<Real*>(p.x_levels);
    const Real eps = (Real)D.eps;
    const int n = D.n, cin = D.c_in;
    c.n = n; c.cin = cin; c.NR0 = n + cin; c.NR = c.NR0; c.z0 = 0;

    V2 t[RPL][NP];
    Real lo[RPL], hi[RPL], yv[RPL], rq[RPL];
    int wh[RPL];

    /* ---- stage the instance blob: one TMA bulk copy; initialise the basis meanwhile ---- */
    blob_stage(c.blob, static_cast<const Real*>(p.blobs) + (size_t)inst * stride_e,
               (uint32_t)stride_e * (uint32_t)sizeof(Real), ws + W.off_bar, lane);
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        if (row < W.NRmax) {
            V2 zero; zero.x = Real(0.0); zero.y = Real(0.0);
            V2* r2 = reinterpret_cast<V2*>(c.ZB + row * LD);
#pragma unroll
            for (int k = 0; k < NP; ++k) r2[k] = zero;
            if (row < n) c.ZB[row * LD + row] = Real(1.0);
            c.yprev[row] = Real(0.0);
        }
    }
    blob_wait(ws + W.off_bar);
    const Real* blob = c.blob;

    /* ---- bounds: l = max l_s, u = min u_s, clipped to +-1e20; global rows lA <= Cin x <= uA ---- */
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        Real l = -Real(OSB_INFTY), h = Real(OSB_INFTY);
        if (row < n) {
            for (int b = 0; b < D.n_bound_sets; ++b) {
                l = fmax(l, blob[L.l[b] + row]);
                h = fmin(h, blob[L.u[b] + row]);
            }
        } else if (row < n + cin) {
            l = blob[L.lA + row - n]; h = blob[L.uA + row - n];
        }
        lo[s] = clip_inf<Real>(l); hi[s] = clip_inf<Real>(h);
        yv[s] = Real(0.0); rq[s] = Real(0.0); wh[s] = -1;
    }
#pragma unroll 1
    for (int r = 0; r < cin; ++r)
#pragma unroll 1
        for (int j = lane; j < n; j += 32) c.ZB[(n + r) * LD + j] = blob[L.Cin + r * n + j];

    int status = OSB_OK, total_iters = 0;
    {   /* a NaN anywhere in the instance poisons everything downstream: flag it and stop */
        bool bad = false;
#pragma unroll 1
        for (int i = lane; i < L.stride; i += 32) { const Real v = blob[i]; bad |= (v != v); }
        if (__any_sync(FULL, bad)) status = OSB_NAN_INPUT;
    }
    if (status != OSB_OK) {
#pragma unroll 1
        for (int i = lane; i < n; i += 32) out_x[(size_t)inst * n + i] = Real(0.0);
        if (lane == 0) { p.status[inst] = status; if (p.iters) p.iters[inst] = 0; }
        return;
    }
    __syncwarp();
    bool x_zero = true;   /* x_prev == 0 so far: residuals are the right-hand sides */

    /* ---- global equalities Ceq x = deq: eliminated before any level ---- */
    if (D.c_eq > 0) {
        const int ce = D.c_eq;
        rr_repack<Real, LD>(c, blob + L.Ceq, n, ce, c.M);
        __syncwarp();
        const int rank = rr_sweep<Real, NP, RPL>(c, t, ce);
        /* particular solution: forward substitution on the pivot rows, consistency check on
         * the dependent ones (all lanes redundantly, ce is small) */
        Real* gz = c.pv;
        int npiv = 0;
#pragma unroll 1
        for (int r = 0; r < ce; ++r) {
            Real s = blob[L.deq + r];
            const int pc = c.piv[r];
            const int upto = (pc >= 0) ? pc : npiv;
            if (pc >= 0) ++npiv;
            for (int k = 0; k < upto; ++k) s -= c.M[r * LD + k] * gz[k];
            if (pc >= 0) { if (lane == 0) gz[pc] = s / c.M[r * LD + pc]; }
            else if (fabs(s) > Tr<Real>::eq_tol * fmax(Real(1.0), fabs(blob[L.deq + r]))) status = OSB_INFEASIBLE;
            __syncwarp();
        }
#pragma unroll 1
        for (int i = lane; i < c.NR0; i += 32) c.yprev[i] = rr_dot_dyn<Real>(c.ZB + i * LD, gz, rank);
        c.z0 = rank;
        x_zero = false;
        __syncwarp();
    }

    int last_level = -1;
    for (int k = 0; k < D.n_levels; ++k) if (p.active_mask & (1u << k)) last_level = k;

#pragma unroll 1
    for (int k = 0; k < D.n_levels && status == OSB_OK; ++k) {
        if (!(p.active_mask & (1u << k))) continue;
        const int wk = D.weight_kind[k];
        const int m = D.m[k];
        const int cl = D.c_loc[k];
        const int z0 = c.z0, nz = n - z0;
        c.NR = c.NR0 + cl;
        const int NR = c.NR;
        const int max_iter = p.max_iter > 0 ? p.max_iter : 10 * (NR + n) + 20;

        if (nz == 0) { /* nothing left to decide: the solution is inherited */
            if (p.store_levels)
                for (int i = lane; i < n; i += 32)
                    out_xl[((size_t)inst * D.n_levels + k) * n + i] = c.yprev[i];
            continue;
        }
        /* ---- local rows: limits and value at x_prev ---- */
        if (cl > 0) {
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int row = lane + 32 * s;
                if (row >= c.NR0) {
                    const bool in = row < NR;
                    lo[s] = in ? clip_inf<Real>(blob[L.lCl[k] + row - c.NR0]) : -Real(OSB_INFTY);
                    hi[s] = in ? clip_inf<Real>(blob[L.uCl[k] + row - c.NR0]) : Real(OSB_INFTY);
                }
            }
            rr_residual<Real>(c, blob + L.Cl[k], n, nullptr, cl, c.yprev, c.yprev + c.NR0);
#pragma unroll 1
            for (int r = lane; r < cl; r += 32) c.yprev[c.NR0 + r] = -c.yprev[c.NR0 + r];
        } else {
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int row = lane + 32 * s;
                if (row >= c.NR0) { lo[s] = -Real(OSB_INFTY); hi[s] = Real(OSB_INFTY); }
            }
        }
        /*
         * Two ways to factor the reduced Hessian Hz = F F' of the level (DENSE-path stacks are
         * solved by sot_kernel.cuh):
         *  TRIVIAL  A = I, W = I:  Hz = (1+eps) I, nothing to factor.
         *  LQ       dense A, diagonal W: rotate the basis FIRST with the Householder sweep of
         *           M = A Z  (M Q = [Lm 0]); in the rotated basis Hz = blkdiag(Lm'W Lm + eps I,
         *           eps I): only a rank x rank Cholesky, and the same sweep is the basis
         *           shrink the next level needs.
         * p (c.pv, absolute columns) becomes F^-1 g_z, the rows of ZB become N = (E Z) F^-T.
         */
        const int path = sot_level_path_of(&D, k);
        int ncol = 0;
        Real tail_scale;
        if (path == SOT_PATH_TRIVIAL) {
            tail_scale = Real(1.0) / sqrt(Real(1.0) + eps);
#pragma unroll 1
            for (int i = lane; i < n; i += 32)
                c.res[i] = (Real(1.0) + eps) * c.yprev[i] - blob[L.b[k] + i] + (D.has_c[k] ? blob[L.c[k] + i] : Real(0.0));
            __syncwarp();
#pragma unroll 1
            for (int j = lane; j < LD; j += 32) {
                Real s0 = 0, s1 = 0;
                if (j >= z0 && j < n) {
                    const Real* zc = c.ZB + j;
                    int i = 0;
#pragma unroll 4
                    for (; i + 1 < n; i += 2) { s0 = fma(c.res[i], zc[i * LD], s0); s1 = fma(c.res[i + 1], zc[(i + 1) * LD], s1); }
                    if (i < n) s0 = fma(c.res[i], zc[i * LD], s0);
                }
                c.pv[j] = s0 + s1;
            }
        } else {
            if (z0 == 0 && D.c_eq == 0) rr_repack<Real, LD>(c, blob + L.A[k], n, m, c.M); /* basis still the identity: M = A */
            else rr_project<Real, LD>(c, blob + L.A[k], n, m, c.M);
            if (x_zero) {
#pragma unroll 1
                for (int r = lane; r < m; r += 32) c.res[r] = blob[L.b[k] + r];
            } else {
                rr_residual<Real>(c, blob + L.A[k], n, blob + L.b[k], m, c.yprev, c.res);
            }
            __syncwarp();
            const int rank = rr_sweep<Real, NP, RPL>(c, t, m);
            ncol = rank; tail_scale = Real(1.0) / sqrt(eps);
            const Real* wd = (wk == OSB_W_DIAG) ? blob + L.w[k] : nullptr;
            /* ---- G = Lm' W Lm + eps I: lane a owns row a (registers), stored to V + a*LD ---- */
            {
                Real* G = c.V;
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    if (32 * s < ncol) {
                        const int a = lane + 32 * s;
                        rr_zero<Real, NP>(t[s]);
                        if (a < ncol) {
#pragma unroll 1
                            for (int r = 0; r < m; ++r) {
                                const Real sc = c.M[r * LD + z0 + a] * (wd ? wd[r] : Real(1.0));
                                const V2* m2 = reinterpret_cast<const V2*>(c.M + r * LD);
#pragma unroll
                                for (int kk = 0; kk < NP; ++kk) {
                                    const V2 mv = m2[kk];
                                    t[s][kk].x = fma(sc, mv.x, t[s][kk].x); t[s][kk].y = fma(sc, mv.y, t[s][kk].y);
                                }
                            }
                        }
                    }
                }
                __syncwarp();   /* the reflectors in V are dead: G takes their place */
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int a = lane + 32 * s;
                    if (32 * s < ncol && a < ncol) {
                        rr_store<Real, NP>(G + a * LD, t[s]);
                        G[a * LD + z0 + a] += eps;
                    }
                }
            }
            /* ---- g_z = eps Z' x_prev (+ Z' c) - Lm' W res ---- */
#pragma unroll 1
            for (int r = lane; r < m; r += 32) c.res[r] *= (wd ? wd[r] : Real(1.0));
            __syncwarp();
#pragma unroll 1
            for (int j = lane; j < LD; j += 32) {
                Real s0 = 0, s1 = 0;
                if (j >= z0 && j < n) {
                    const Real* zc = c.ZB + j;
                    if (!x_zero || D.has_c[k]) {
#pragma unroll 4
                        for (int i = 0; i < n; ++i) {
                            const Real vi = eps * c.yprev[i] + (D.has_c[k] ? blob[L.c[k] + i] : Real(0.0));
                            s0 = fma(vi, zc[i * LD], s0);
                        }
                    }
                    if (j < z0 + ncol) {
                        const Real* mc = c.M + j;
#pragma unroll 4
                        for (int r = 0; r < m; ++r) s1 = fma(c.res[r], mc[r * LD], s1);
                    }
                }
                c.pv[j] = s0 - s1;
            }
        }
        __syncwarp();
        /* ---- local rows in the (rotated) basis: extra rows of ZB ---- */
        if (cl > 0) { rr_project<Real, LD>(c, blob + L.Cl[k], n, cl, c.ZB + c.NR0 * LD); __syncwarp(); }
        /*
         * ---- one left-looking pass over the leading columns does three jobs at once: the
         * Cholesky factor of G (rows of G), p = F^-1 g_z (the gradient as one more row) and the
         * leading columns of N = (E Z) F^-T (the basis rows, in place: those columns leave
         * the basis after this level). Row X: X[c] = (X[c] - X[:c] . L[c][:c]) / L[c][c]. ----
         */
        bool spd = true;
        {
            const Real* G = c.V;
            const int gs_n = (ncol + 1 + 31) >> 5;        /* passes over [G rows ; gradient row] */
#pragma unroll 1
            for (int col = 0; col < ncol; ++col) {
                const Real* Lc = G + col * LD + z0;
                Real sb[RPL], sg0 = 0, sg1 = 0;
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int row = lane + 32 * s;
                    sb[s] = Real(0.0);
                    if (row < NR) {
                        const Real* X = c.ZB + row * LD + z0;
                        sb[s] = X[col] - rr_dot_dyn<Real>(X, Lc, col);
                    }
                }
                {
                    const int g = lane;
                    if (g >= col && g <= ncol) {
                        const Real* X = (g < ncol) ? (G + g * LD + z0) : (c.pv + z0);
                        sg0 = X[col] - rr_dot_dyn<Real>(X, Lc, col);
                    }
                    if (gs_n > 1) {
                        const int g1 = lane + 32;
                        if (g1 >= col && g1 <= ncol) {
                            const Real* X = (g1 < ncol) ? (G + g1 * LD + z0) : (c.pv + z0);
                            sg1 = X[col] - rr_dot_dyn<Real>(X, Lc, col);
                        }
                    }
                }
                const Real piv = __shfl_sync(FULL, (col < 32) ? sg0 : sg1, col & 31);
                if (!(piv > Real(0.0))) { spd = false; break; }
                const Real inv = rsqrt(piv);
                __syncwarp();
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    const int row = lane + 32 * s;
                    if (row < NR) c.ZB[row * LD + z0 + col] = sb[s] * inv;
                }
                {
                    const int g = lane;
                    if (g >= col && g <= ncol) {
                        Real* X = (g < ncol) ? (c.V + g * LD + z0) : (c.pv + z0);
                        X[col] = sg0 * inv;
                    }
                    const int g1 = lane + 32;
                    if (gs_n > 1 && g1 >= col && g1 <= ncol) {
                        Real* X = (g1 < ncol) ? (c.V + g1 * LD + z0) : (c.pv + z0);
                        X[col] = sg1 * inv;
                    }
                }
                __syncwarp();
            }
        }
        if (!spd) { status = OSB_NOT_SPD | (k << 8); break; }
        /* ---- column scales of the remaining columns, p on the tail ---- */
#pragma unroll 1
        for (int j = lane; j < LD; j += 32) {
            const bool live = (j >= z0) && (j < n);
            const bool lead = live && (j < z0 + ncol);
            c.cscale[j] = lead ? Real(1.0) : (live ? tail_scale : Real(0.0));
            if (!lead) c.pv[j] = live ? c.pv[j] * tail_scale : Real(0.0);
        }
        __syncwarp();
        /* ---- rows of N into registers; unconstrained minimiser y = y_prev - N p; |row|^2 ---- */
        {
            const V2* cs2 = reinterpret_cast<const V2*>(c.cscale);
            const V2* p2 = reinterpret_cast<const V2*>(c.pv);
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int row = lane + 32 * s;
                if (row < NR) {
                    const V2* z2 = reinterpret_cast<const V2*>(c.ZB + row * LD);
                    Real a0 = 0, a1 = 0, q0 = 0, q1 = 0;
#pragma unroll
                    for (int kk = 0; kk < NP; ++kk) {
                        const V2 zv = z2[kk], sc = cs2[kk], pp = p2[kk];
                        V2 tv; tv.x = zv.x * sc.x; tv.y = zv.y * sc.y;
                        t[s][kk] = tv;
                        a0 = fma(tv.x, pp.x, a0); a1 = fma(tv.y, pp.y, a1);
                        q0 = fma(tv.x, tv.x, q0); q1 = fma(tv.y, tv.y, q1);
                    }
                    yv[s] = c.yprev[row] - (a0 + a1);
                    rq[s] = q0 + q1;     /* |row|^2 of N: invariant under the active-set column rotations */
                } else {
                    rr_zero<Real, NP>(t[s]);
                    yv[s] = Real(0.0); rq[s] = Real(0.0);
                }
            }
        }
        __syncwarp();
        /* ---- inequality rows and bounds ---- */
        RRLevelOut lo_out = rr_active_set<Real, NP, RPL>(c, t, yv, lo, hi, rq, wh, max_iter);
        total_iters += lo_out.iters;
        if (lo_out.status != OSB_OK) { status = lo_out.status | (k << 8); break; }
        /* ---- level solution; active bounds are returned exactly on the bound ---- */
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int row = lane + 32 * s;
            if (row < c.NR0) {
                Real v = yv[s];
                if (row < n && wh[s] >= 0) v = (c.act[wh[s]] & 0x10000) ? hi[s] : lo[s];
                c.yprev[row] = v;
                if (p.store_levels && row < n) out_xl[((size_t)inst * D.n_levels + k) * n + row] = v;
            }
        }
        x_zero = false;
        __syncwarp();
        /* ---- shrink the basis by the row space of this level's task ---- */
        if (path == SOT_PATH_LQ) c.z0 += ncol;
        else if (k != last_level) c.z0 = n;
    }
    /* ---- result ---- */
#pragma unroll 1
    for (int i = lane; i < n; i += 32) out_x[(size_t)inst * n + i] = c.yprev[i];
    if (lane == 0) {
        p.status[inst] = status;
        if (p.iters) p.iters[inst] = total_iters;
    }
}

#ifdef OSB_EMU
template <typename Real>
__device__ void rr_dispatch(const sot_params& p)
{
    if (p.rr.np == 15 && p.rr.rpl == 1) rr_solve_impl<Real, 15, 1>(p);
    else if (p.rr.np == 15 && p.rr.rpl == 2) rr_solve_impl<Real, 15, 2>(p);
    else rr_solve_impl<Real, 19, 2>(p);
}
#endif

} // namespace

#ifndef OSB_EMU
/* launch bounds: resident warps per SM the register budget must allow (shared memory is the
 * other limit; rr_plan_workspace sizes it) */
template <typename Real, int NP, int RPL, int MINB>
__global__ void __launch_bounds__(32, MINB) rr_solve_kernel(const __grid_constant__ sot_params p)
{
    rr_solve_impl<Real, NP, RPL>(p);
}
#endif

#endif

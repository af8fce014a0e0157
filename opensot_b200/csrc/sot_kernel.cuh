/*
 * sot_kernel.cuh -- the hot path: one warp solves one whole stack of tasks.
 *
 * Replaces, per instance and per control cycle:
 *   a-2  bounds intersection / +-1e20 clipping   src/constraints/Aggregated.cpp:141-148,
 *                                                src/solvers/QPOasesBackEnd.cpp:339-356
 *   a-5  H = A'WA + eps I, g = -A'Wb + c         src/solvers/iHQP.cpp:129-162,274-278
 *                                                src/solvers/QPOasesBackEnd.cpp:253-255
 *   a-6  optimality rows A_j x = A_j x_j*        src/solvers/iHQP.cpp:164-170,285-332
 *   a-8  the dense QP of every level             src/solvers/QPOasesBackEnd.cpp:248-307
 *   a-9  the level loop / active levels          src/solvers/iHQP.cpp:263-358
 *
 * Formulation (not the reference's): equalities are never handed
 * to the QP. After level j the feasible set of later levels is x_j* + null(A_0..A_j), so
 * the warp carries an orthonormal basis Z of that null space (Householder reflections
 * accumulated into ZB = [Z ; Cin Z]) and solves level k in z with x = x_prev + Z z:
 *     H_z = (A_k Z)' W (A_k Z) + eps I ,  g_z = -(A_k Z)' W (b_k - A_k x_prev) + eps Z' x_prev
 * which is exactly the reference's level-k QP (eps |x|^2 kept on the FULL variable)
 * restricted to the set the optimality equalities describe. Bounds and inequality rows are
 * box constraints on y = [x ; Cin x ; Cloc x] = y_prev + (E Z) z and are handled by a
 * Goldfarb-Idnani dual active-set iteration on N = (E Z) L^-T (H_z = L L'), with
 * Householder appends and Givens removals of working-set columns.
 *
 * Mapping: CTA = 1 warp = 1 instance (grid = batch; the hardware block scheduler does the
 * load balancing between instances with different active-set iteration counts). The
 * instance blob is staged with one cp.async.bulk (TMA bulk copy, UBLKCP) + mbarrier. All
 * matrices live in shared memory, ROW-major with one odd row stride LD: the inner loops of
 * every phase run along a lane's own row (consecutive words, immediate-offset LDS once
 * unrolled) against a broadcast vector, and column-wise accesses of a warp hit 32 banks.
 */
#ifndef OSB_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include "sot_params.h"
#include "sot_common.cuh"

#ifdef OSB_PROFILE
#define PROF_T(var) const long long var = clock64()
#define PROF_PRINT(...) do { if (threadIdx.x == 0 && blockIdx.x == OSB_PROFILE) printf(__VA_ARGS__); } while (0)
#else
#define PROF_T(var)
#define PROF_PRINT(...)
#endif

namespace {

/* ---- contiguous primitives: the inner loops of every phase -------------------------- */
/*
 * Rows are 16-byte aligned (even row stride), so a lane walks its row with LDS.128 (two
 * doubles per load; with LD/2 odd the eight lanes of a quarter-warp hit eight different
 * 16-byte bank groups). Both operands must be in the same 16-byte phase for the vector
 * path; otherwise (only cold call sites) a scalar path is taken.
 */
template <typename Real>
__device__ __forceinline__ bool same_phase(const Real* a, const Real* b)
{
    return (((uintptr_t)a ^ (uintptr_t)b) & sizeof(Real)) == 0;
}
/* aligned pair walks: a2/b2 point at 16-byte aligned doubles, n2 = number of typename Tr<Real>::V2 pairs.
 * Bodies of four pairs with every load issued before the first FMA: one shared-memory
 * latency (~35-70 cycles on B200, measured) is paid per eight elements, not per two.
 * (Bodies that mask a clamped surplus instead of a remainder loop were measured slower: the
 * extra loads cost more shared-memory bandwidth than the branches they save.) */
template <typename Real>
__device__ __forceinline__ Real dot_pairs(const typename Tr<Real>::V2* __restrict__ a2, const typename Tr<Real>::V2* __restrict__ b2, int n2)
{
    Real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int j = 0;
#pragma unroll 1
    for (; j + 3 < n2; j += 4) {
        const typename Tr<Real>::V2 x0 = a2[j], x1 = a2[j + 1], x2 = a2[j + 2], x3 = a2[j + 3];
        const typename Tr<Real>::V2 y0 = b2[j], y1 = b2[j + 1], y2 = b2[j + 2], y3 = b2[j + 3];
        s0 = fma(x0.x, y0.x, s0); s1 = fma(x0.y, y0.y, s1); s2 = fma(x1.x, y1.x, s2); s3 = fma(x1.y, y1.y, s3);
        s0 = fma(x2.x, y2.x, s0); s1 = fma(x2.y, y2.y, s1); s2 = fma(x3.x, y3.x, s2); s3 = fma(x3.y, y3.y, s3);
    }
    if (j + 1 < n2) {
        const typename Tr<Real>::V2 x0 = a2[j], x1 = a2[j + 1], y0 = b2[j], y1 = b2[j + 1];
        s0 = fma(x0.x, y0.x, s0); s1 = fma(x0.y, y0.y, s1); s2 = fma(x1.x, y1.x, s2); s3 = fma(x1.y, y1.y, s3);
        j += 2;
    }
    if (j < n2) {
        const typename Tr<Real>::V2 x0 = a2[j], y0 = b2[j];
        s0 = fma(x0.x, y0.x, s0); s1 = fma(x0.y, y0.y, s1);
    }
    return (s0 + s1) + (s2 + s3);
}
template <typename Real>
__device__ __forceinline__ void axpy_pairs(typename Tr<Real>::V2* __restrict__ y2, const typename Tr<Real>::V2* __restrict__ x2, Real w, int n2)
{
    int j = 0;
#pragma unroll 1
    for (; j + 3 < n2; j += 4) {
        typename Tr<Real>::V2 a0 = y2[j], a1 = y2[j + 1], a2 = y2[j + 2], a3 = y2[j + 3];
        const typename Tr<Real>::V2 b0 = x2[j], b1 = x2[j + 1], b2 = x2[j + 2], b3 = x2[j + 3];
        a0.x = fma(-w, b0.x, a0.x); a0.y = fma(-w, b0.y, a0.y); a1.x = fma(-w, b1.x, a1.x); a1.y = fma(-w, b1.y, a1.y);
        a2.x = fma(-w, b2.x, a2.x); a2.y = fma(-w, b2.y, a2.y); a3.x = fma(-w, b3.x, a3.x); a3.y = fma(-w, b3.y, a3.y);
        y2[j] = a0; y2[j + 1] = a1; y2[j + 2] = a2; y2[j + 3] = a3;
    }
    if (j + 1 < n2) {
        typename Tr<Real>::V2 a0 = y2[j], a1 = y2[j + 1];
        const typename Tr<Real>::V2 b0 = x2[j], b1 = x2[j + 1];
        a0.x = fma(-w, b0.x, a0.x); a0.y = fma(-w, b0.y, a0.y); a1.x = fma(-w, b1.x, a1.x); a1.y = fma(-w, b1.y, a1.y);
        y2[j] = a0; y2[j + 1] = a1;
        j += 2;
    }
    if (j < n2) {
        typename Tr<Real>::V2 a0 = y2[j];
        const typename Tr<Real>::V2 b0 = x2[j];
        a0.x = fma(-w, b0.x, a0.x); a0.y = fma(-w, b0.y, a0.y);
        y2[j] = a0;
    }
}
/* sum_j a[j] * b[j] over an arbitrary range: peel to alignment, then pair walk */
template <typename Real>
__device__ __forceinline__ Real dot_cc(const Real* __restrict__ a, const Real* __restrict__ b, int cnt)
{
    Real s = 0;
    if (same_phase(a, b)) {
        if (((uintptr_t)a & sizeof(Real)) && cnt > 0) { s = a[0] * b[0]; ++a; ++b; --cnt; }
        s += dot_pairs<Real>(reinterpret_cast<const typename Tr<Real>::V2*>(a), reinterpret_cast<const typename Tr<Real>::V2*>(b), cnt >> 1);
        if (cnt & 1) s = fma(a[cnt - 1], b[cnt - 1], s);
    } else {
        Real s1 = 0;
        int j = 0;
#pragma unroll 1
        for (; j + 1 < cnt; j += 2) { s = fma(a[j], b[j], s); s1 = fma(a[j + 1], b[j + 1], s1); }
        if (j < cnt) s = fma(a[j], b[j], s);
        s += s1;
    }
    return s;
}
/* y[j] -= w * x[j] */
template <typename Real>
__device__ __forceinline__ void axpy_cc(Real* __restrict__ y, const Real* __restrict__ x, Real w, int cnt)
{
    if (same_phase(y, x)) {
        if (((uintptr_t)y & sizeof(Real)) && cnt > 0) { y[0] = fma(-w, x[0], y[0]); ++y; ++x; --cnt; }
        axpy_pairs<Real>(reinterpret_cast<typename Tr<Real>::V2*>(y), reinterpret_cast<const typename Tr<Real>::V2*>(x), w, cnt >> 1);
        if (cnt & 1) y[cnt - 1] = fma(-w, x[cnt - 1], y[cnt - 1]);
    } else {
#pragma unroll 1
        for (int j = 0; j < cnt; ++j) y[j] = fma(-w, x[j], y[j]);
    }
}
/* sum_i a[i] * b[i * ldb] (b strided: one column of a row-major matrix) */
template <typename Real>
__device__ __forceinline__ Real dot_cs(const Real* __restrict__ a, const Real* __restrict__ b, int ldb, int cnt)
{
    Real s0 = 0, s1 = 0;
    int i = 0;
#pragma unroll 1
    for (; i + 1 < cnt; i += 2) {
        s0 = fma(a[i], b[0], s0);
        s1 = fma(a[i + 1], b[ldb], s1);
        b += 2 * ldb;
    }
    if (i < cnt) s0 = fma(a[i], b[0], s0);
    return s0 + s1;
}

/* Everything one warp needs, as shared-memory pointers. */
template <typename Real>
struct Ctx {
    const sot_params* p;
    int lane;
    Real *blob, *ZB, *T, *R, *X;
    Real *yprev, *y, *ylo, *yhi, *dy, *d, *dm, *rowsq, *r, *u, *rinv, *gz, *res, *rn2;
    int32_t *act, *where, *piv;
    int LD;
    int n, cin, NR0;    /* NR0 = n + cin : rows carried from level to level               */
    int z0, nz;         /* current basis = columns [z0, z0+nz) of ZB                      */
    int NR;             /* rows of y at this level (NR0 + c_loc)                          */
    int mo;             /* z0 & 1: word offset of basis column 0 inside T rows, d, dm, gz  */
    int q;
};

/* prefix dot of two 16-byte aligned rows: sum_{j<cnt} a[j] b[j] */
template <typename Real>
__device__ __forceinline__ Real dot_prefix(const Real* __restrict__ a, const Real* __restrict__ b, int cnt)
{
    Real s = dot_pairs<Real>(reinterpret_cast<const typename Tr<Real>::V2*>(a), reinterpret_cast<const typename Tr<Real>::V2*>(b), cnt >> 1);
    if (cnt & 1) s = fma(a[cnt - 1], b[cnt - 1], s);
    return s;
}

/*
 * Householder sweep that pushes the row space of `m` rows M (row r at Mb + r*LD, in the
 * coordinates of the current basis) out of the basis: afterwards those rows are [L 0] and
 * the trailing nz - rank columns of ZB span the null space.
 * Replaces adding the optimality equalities of a finished level to every later QP
 * (iHQP.cpp:164-170, 314-332). Linearly dependent rows (|residual| <= tol_rank * |row|)
 * are skipped, the job qpOASES' linear-independence test does (QProblem.cpp:2847).
 * Returns the rank; c.piv[r] = pivot column of row r or -1.
 *
 * Two phases, so that the long part has no barriers and no scalar recurrences:
 *  1. LQ of M alone (lane per remaining task row). Reflector k (H = I - tau v v', v[pc] = 1)
 *     is written as a full, zero-filled row V_k into T (free at every call site), tau in
 *     c.gz[k], so that every later use is a branch-free walk of aligned 128-bit pairs;
 *  2. every basis row (lane per row) applies V_0 .. V_{rank-1} back to back.
 * Mb must share the 16-byte phase of the basis rows: Mb = row start + (z0 & 1); V rows are
 * shifted the same way, and walks start on the even word at or before the pivot (the word
 * before the pivot belongs to a finished column and meets a zero of V).
 */
template <typename Real>
__device__ int nullspace_sweep(Ctx<Real>& c, Real* Mb, int m)
{
    const int lane = c.lane, LD = c.LD, nz = c.nz;
    const int mo = c.z0 & 1;
    const Real tol2 = (Real)(c.p->tol_rank * c.p->tol_rank);
#pragma unroll 1
    for (int r = lane; r < m; r += 32) { /* squared norm of every row before the sweep */
        Real* row = Mb + r * LD;
        c.rn2[r] = dot_cc<Real>(row, row, nz);
        c.piv[r] = -1;
        if (mo) row[-1] = Real(0.0);   /* guard words met by the aligned pair walks */
        if ((mo + nz) & 1) row[nz] = Real(0.0);
    }
    __syncwarp();
    int zc = 0;
#pragma unroll 1
    for (int r = 0; r < m && zc < nz; ++r) {
        Real* Mr = Mb + r * LD;
        Real part = 0;
#pragma unroll 1
        for (int j = zc + lane; j < nz; j += 32) { const Real a = Mr[j]; part = fma(a, a, part); }
        const Real nrm2 = warp_sum<Real>(part);
        const Real a_zc = Mr[zc];
        if (!(nrm2 > tol2 * c.rn2[r]) || nrm2 <= Real(0.0)) { /* dependent (or zero) row: drop its noise */
            __syncwarp();
#pragma unroll 1
            for (int j = zc + lane; j < nz; j += 32) Mr[j] = Real(0.0);
            __syncwarp();
            continue;
        }
        const Real rs = rsqrt(nrm2), sq = nrm2 * rs;
        const Real rho = (a_zc > Real(0.0)) ? -sq : sq;
        const Real rho_inv = (a_zc > Real(0.0)) ? -rs : rs;
        const Real scale = Tr<Real>::rcp(a_zc - rho);   /* v = (a - rho e1) / (a_zc - rho) */
        const Real tau = (rho - a_zc) * rho_inv;     /* H = I - tau v v'                */
        Real* V = c.T + zc * LD;                     /* reflector zc, phase-shifted like Mb */
#pragma unroll 1
        for (int jj = lane; jj < LD; jj += 32) {
            const int j = jj - mo;                     /* basis column of word jj */
            Real vj = Real(0.0);
            if (j == zc) vj = Real(1.0);
            else if (j > zc && j < nz) vj = Mr[j] * scale;
            V[jj] = vj;
        }
        __syncwarp();
#pragma unroll 1
        for (int j = zc + lane; j < nz; j += 32) Mr[j] = (j == zc) ? rho : Real(0.0); /* row r is now [L 0] */
        if (lane == 0) { c.gz[zc] = tau; c.piv[r] = zc; }
        /* remaining task rows, lane per row; walk from the even word at/before the pivot */
        const int e0 = (mo + zc) & ~1;                 /* word offset inside a row */
        const int n2 = (mo + nz + 1 - e0) >> 1;
#pragma unroll 1
        for (int i = r + 1 + lane; i < m; i += 32) {
            typename Tr<Real>::V2* row = reinterpret_cast<typename Tr<Real>::V2*>(Mb - mo + i * LD + e0);
            const typename Tr<Real>::V2* v2 = reinterpret_cast<const typename Tr<Real>::V2*>(V + e0);
            const Real w = tau * dot_pairs<Real>(row, v2, n2);
            axpy_pairs<Real>(row, v2, w, n2);
        }
        __syncwarp();
        ++zc;
    }
    __syncwarp();
    const int rank = zc;
    /* phase 2: basis rows (x rows and Cin rows) take all reflectors, lane per row */
#pragma unroll 1
    for (int i = lane; i < c.NR0; i += 32) {
        Real* rowb = c.ZB + i * LD + c.z0 - mo;      /* word 0 of the phase-shifted view */
#pragma unroll 1
        for (int k = 0; k < rank; ++k) {
            const int e0 = (mo + k) & ~1;
            const int n2 = (mo + nz + 1 - e0) >> 1;
            typename Tr<Real>::V2* row = reinterpret_cast<typename Tr<Real>::V2*>(rowb + e0);
            const typename Tr<Real>::V2* v2 = reinterpret_cast<const typename Tr<Real>::V2*>(c.T + k * LD + e0);
            const Real w = c.gz[k] * dot_pairs<Real>(row, v2, n2);
            axpy_pairs<Real>(row, v2, w, n2);
        }
    }
    __syncwarp();
    return rank;
}

/* dst rows (row r at dst + r*LD)  <-  A (m x n row-major, row stride lda) times the basis;
 * lane per basis column; a body takes four rows of A times four basis words, all 20 loads
 * issued before the 16 FMAs */
template <typename Real>
__device__ void project_rows(Ctx<Real>& c, const Real* __restrict__ A, int lda, int m, Real* dst)
{
    const int LD = c.LD, n = c.n;
#pragma unroll 1
    for (int j = c.lane; j < c.nz; j += 32) {
        const Real* zcol = c.ZB + c.z0 + j;
        int r = 0;
#pragma unroll 1
        for (; r + 3 < m; r += 4) {
            const Real *a0 = A + r * lda, *a1 = a0 + lda, *a2 = a1 + lda, *a3 = a2 + lda;
            const Real* z = zcol;
            Real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int i = 0;
#pragma unroll 1
            for (; i + 3 < n; i += 4, z += 4 * LD) {
                const Real z0 = z[0], z1 = z[LD], z2 = z[2 * LD], z3 = z[3 * LD];
                const Real p0 = a0[i], p1 = a0[i + 1], p2 = a0[i + 2], p3 = a0[i + 3];
                const Real q0 = a1[i], q1 = a1[i + 1], q2 = a1[i + 2], q3 = a1[i + 3];
                const Real u0 = a2[i], u1 = a2[i + 1], u2 = a2[i + 2], u3 = a2[i + 3];
                const Real v0 = a3[i], v1 = a3[i + 1], v2 = a3[i + 2], v3 = a3[i + 3];
                s0 = fma(p0, z0, s0); s1 = fma(q0, z0, s1); s2 = fma(u0, z0, s2); s3 = fma(v0, z0, s3);
                s0 = fma(p1, z1, s0); s1 = fma(q1, z1, s1); s2 = fma(u1, z1, s2); s3 = fma(v1, z1, s3);
                s0 = fma(p2, z2, s0); s1 = fma(q2, z2, s1); s2 = fma(u2, z2, s2); s3 = fma(v2, z2, s3);
                s0 = fma(p3, z3, s0); s1 = fma(q3, z3, s1); s2 = fma(u3, z3, s2); s3 = fma(v3, z3, s3);
            }
#pragma unroll 1
            for (; i < n; ++i, z += LD) {
                const Real zv = *z;
                s0 = fma(a0[i], zv, s0); s1 = fma(a1[i], zv, s1);
                s2 = fma(a2[i], zv, s2); s3 = fma(a3[i], zv, s3);
            }
            dst[r * LD + j] = s0; dst[(r + 1) * LD + j] = s1;
            dst[(r + 2) * LD + j] = s2; dst[(r + 3) * LD + j] = s3;
        }
#pragma unroll 1
        for (; r < m; ++r) dst[r * LD + j] = dot_cs<Real>(A + r * lda, zcol, LD, n);
    }
}
/* out[r] = rhs[r] - A[r,:] . x   (lane per row) */
template <typename Real>
__device__ void residual_rows(Ctx<Real>& c, const Real* __restrict__ A, int lda, const Real* rhs, int m,
                              const Real* x, Real* out)
{
#pragma unroll 1
    for (int r = c.lane; r < m; r += 32)
        out[r] = (rhs ? rhs[r] : Real(0.0)) - dot_cc<Real>(A + r * lda, x, c.n);
}
/* lower triangle of R (row-major L[a][b], b <= a < ncols) += P' diag(w) Q over `rows` rows,
 * P, Q with row stride LD; one (a,b) pair per lane and pass */
template <typename Real>
__device__ void gram_pairs(Ctx<Real>& c, Real* G, const Real* P, const Real* Q, int rows, const Real* wd, int ncols)
{
    const int LD = c.LD;
    const int npairs = ncols * (ncols + 1) / 2;
#pragma unroll 1
    for (int idx = c.lane; idx < npairs; idx += 32) {
        int a = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while ((a + 1) * (a + 2) / 2 <= idx) ++a;
        while (a * (a + 1) / 2 > idx) --a;
        const int b = idx - a * (a + 1) / 2;
        const Real *pa = P + a, *qb = Q + b;
        Real s0 = 0, s1 = 0;
        int r = 0;
#pragma unroll 1
        for (; r + 1 < rows; r += 2) {
            s0 = fma(pa[r * LD] * (wd ? wd[r] : Real(1.0)), qb[r * LD], s0);
            s1 = fma(pa[(r + 1) * LD] * (wd ? wd[r + 1] : Real(1.0)), qb[(r + 1) * LD], s1);
        }
        if (r < rows) s0 = fma(pa[r * LD] * (wd ? wd[r] : Real(1.0)), qb[r * LD], s0);
        G[a * LD + b] += s0 + s1;
    }
}

struct LevelOut { int status; int iters; };

/*
 * Goldfarb-Idnani dual active-set iteration on the box ylo <= y <= yhi, y = y0 + N z,
 * N = rows [0,NR) of T (already multiplied by L^-T), working set factor R (q x q upper).
 * The single-QP job of BackEnd::solve (QPOasesBackEnd.cpp:248-307); method after
 * Goldfarb & Idnani 1983 as used by the reference's eiQuadProg back-end
 * (external/eiQuadProg-ext/include/eiquadprog.hpp:287-470), re-organised for one warp:
 *  - lane per row for everything that touches N or y; the candidate's row d = N' n_p is
 *    published once per step in shared memory, plain (c.d) and masked to the free columns
 *    (c.dm), so that every row-walk is a branch-free run of aligned 128-bit loads;
 *  - |d2|^2 is accumulated by every lane alongside its own dot product (no shuffles);
 *    |d|^2 is the squared row norm of N, invariant under the orthogonal column updates and
 *    therefore computed once per level (c.rowsq);
 *  - pricing and ratio test end in two 32-bit redux operations each.
 */
template <typename Real>
__device__ LevelOut dual_active_set(Ctx<Real>& c, int max_iter)
{
    const int lane = c.lane, LD = c.LD, nz = c.nz, NR = c.NR;
    const int mo = c.mo;                      /* word of basis column 0 in T rows, d, dm */
    const int wE = (mo + nz + 1) & ~1;        /* even end of the word range of a row walk */
    const Real tol = (Real)c.p->tol_feas;
    LevelOut out{OSB_OK, 0};
    bool parked = false, refreshed = false;
    c.q = 0;
#pragma unroll 1
    for (int i = lane; i < NR; i += 32) c.where[i] = -1;
    __syncwarp();
#pragma unroll 1
    for (;;) {
        /* ---- pricing: most violated row that is neither in the working set nor parked ---- */
        PROF_T(gs);
        Real best = Real(0.0); int bidx = -1;
#pragma unroll 1
        for (int i = lane; i < NR; i += 32) {
            const Real v = c.y[i], lo = c.ylo[i], hi = c.yhi[i];
            Real vl = (lo - v) - tol * fmax(Real(1.0), fabs(lo)); /* > 0: violated beyond tolerance */
            Real vh = (v - hi) - tol * fmax(Real(1.0), fabs(hi));
            vl = (lo > -Real(OSB_INFTY)) ? vl : -Real(1.0);
            vh = (hi < Real(OSB_INFTY)) ? vh : -Real(1.0);
            const bool upper = vh > vl;
            const Real cand = upper ? vh : vl;
            if (c.where[i] == -1 && cand > best) { best = cand; bidx = upper ? (i | 0x10000) : i; }
        }
        PROF_T(g0);
        bidx = warp_argmax_nonneg(best, bidx);
        PROF_T(g1);
        if (!(best > Real(0.0))) break; /* primal feasible: optimal */
        const int ip = bidx & 0xffff;
        const Real sgn = (bidx & 0x10000) ? -Real(1.0) : Real(1.0); /* +1: lower side, -1: upper side */
        const Real lim = (sgn > 0) ? c.ylo[ip] : c.yhi[ip];
        const Real* Tip = c.T + ip * LD;
        const Real dd = c.rowsq[ip];
        Real up = Real(0.0);
#pragma unroll 1
        for (;;) {
            if (++out.iters > max_iter) { out.status = OSB_MAX_ITER; return out; }
            const int q = c.q;
            const int j0 = (mo + q) & ~1;             /* even word at or before the first free column */
            /* d = N' n_p (row ip of N, signed); dm = d restricted to the free columns */
#pragma unroll 1
            for (int w = lane; w < wE; w += 32) {
                const int j = w - mo;
                const Real dj = (j >= 0 && j < nz) ? sgn * Tip[w] : Real(0.0);
                c.d[w] = dj;
                c.dm[w] = (j >= q) ? dj : Real(0.0);
            }
            __syncwarp();
            PROF_T(g2);
            /* dy = N2 d2 for every row, |d2|^2 on the side */
            Real zz = 0;
            {
                const typename Tr<Real>::V2* dv = reinterpret_cast<const typename Tr<Real>::V2*>(c.dm + j0);
                const int n2 = (wE - j0) >> 1;
#pragma unroll 1
                for (int i = lane, pass = 0; pass == 0 || i < NR; i += 32, ++pass) {
                    const typename Tr<Real>::V2* row = reinterpret_cast<const typename Tr<Real>::V2*>(c.T + (i < NR ? i : 0) * LD + j0);
                    Real s0 = 0, s1 = 0, s2 = 0, s3 = 0, z0 = 0, z1 = 0;
                    int j = 0;
#pragma unroll 1
                    for (; j + 3 < n2; j += 4) {
                        const typename Tr<Real>::V2 x0 = row[j], x1 = row[j + 1], x2 = row[j + 2], x3 = row[j + 3];
                        const typename Tr<Real>::V2 v0 = dv[j], v1 = dv[j + 1], v2 = dv[j + 2], v3 = dv[j + 3];
                        s0 = fma(x0.x, v0.x, s0); s1 = fma(x0.y, v0.y, s1); s2 = fma(x1.x, v1.x, s2); s3 = fma(x1.y, v1.y, s3);
                        s0 = fma(x2.x, v2.x, s0); s1 = fma(x2.y, v2.y, s1); s2 = fma(x3.x, v3.x, s2); s3 = fma(x3.y, v3.y, s3);
                        z0 = fma(v0.x, v0.x, z0); z1 = fma(v0.y, v0.y, z1); z0 = fma(v1.x, v1.x, z0); z1 = fma(v1.y, v1.y, z1);
                        z0 = fma(v2.x, v2.x, z0); z1 = fma(v2.y, v2.y, z1); z0 = fma(v3.x, v3.x, z0); z1 = fma(v3.y, v3.y, z1);
                    }
#pragma unroll 1
                    for (; j < n2; ++j) {
                        const typename Tr<Real>::V2 x0 = row[j], v0 = dv[j];
                        s0 = fma(x0.x, v0.x, s0); s1 = fma(x0.y, v0.y, s1);
                        z0 = fma(v0.x, v0.x, z0); z1 = fma(v0.y, v0.y, z1);
                    }
                    if (i < NR) c.dy[i] = (s0 + s1) + (s2 + s3);
                    if (pass == 0) zz = z0 + z1;
                }
            }
            PROF_T(g3);
            /* r = R^-1 d1 (column-oriented back substitution, one value per lane) */
            Real t1 = INFINITY; int lk = -1;
            if (q > 0) {
                Real r0 = (lane < q) ? c.d[mo + lane] : Real(0.0);
                Real r1 = (lane + 32 < q) ? c.d[mo + lane + 32] : Real(0.0);
                /* the lane that owns entry j scales it by 1/R[j][j] (kept in a register) before the
                 * broadcast, and column j-1 of R is fetched while step j is in flight: the chain
                 * per step is one FMA, one multiply and one shuffle */
                const Real ri0 = (lane < q) ? c.rinv[lane] : Real(0.0);
                const Real ri1 = (lane + 32 < q) ? c.rinv[lane + 32] : Real(0.0);
                Real col0 = (lane < q - 1) ? c.R[lane * LD + q - 1] : Real(0.0);
#pragma unroll 1
                for (int j = q - 1; j >= 0; --j) {
                    const Real nxt = (j > 0 && lane < j - 1) ? c.R[lane * LD + j - 1] : Real(0.0);
                    const Real mine = (j < 32) ? r0 * ri0 : r1 * ri1;
                    const Real rj = __shfl_sync(FULL, mine, j & 31);
                    if (lane == (j & 31)) { if (j < 32) r0 = rj; else r1 = rj; }
                    r0 = fma(-col0, rj, r0);
                    if (lane + 32 < j) r1 = fma(-c.R[(lane + 32) * LD + j], rj, r1);
                    col0 = nxt;
                }
                if (lane < q) c.r[lane] = r0;
                if (lane + 32 < q) c.r[lane + 32] = r1;
                /* ratio test over droppable (non-equality) working-set members */
#pragma unroll 1
                for (int k = lane; k < q; k += 32) {
                    const Real rk = (k < 32) ? r0 : r1;
                    const int ak = c.act[k] & 0xffff;
                    if (rk > Real(0.0) && !(c.ylo[ak] == c.yhi[ak])) {
                        const Real ratio = fmax(c.u[k] / rk, Real(0.0)); /* a multiplier is never negative */
                        if (ratio < t1) { t1 = ratio; lk = k; }
                    }
                }
                lk = warp_argmin_nonneg(t1, lk);
            }
            __syncwarp();
            PROF_T(g4);
            const Real s = sgn * (c.y[ip] - lim); /* < 0 while violated */
            const bool indep = (zz > Tr<Real>::dep_tol * dd) && (q < nz);
            const Real t2 = indep ? -s / zz : INFINITY;
            const Real t = fmin(t1, t2);
#ifdef OSB_DEBUG
            if (lane == 0 && blockIdx.x == OSB_DEBUG)
                printf("it %d ip %d sgn %+.0f q %d nz %d s %.6e zz %.6e dd %.6e t1 %.6e lk %d t2 %.6e\n",
                       out.iters, ip, sgn, q, nz, s, zz, dd, t1, lk, t2);
#endif
            if (!(t < INFINITY) && q == nz && !refreshed) {
                /* full vertex: the iterate is determined by the working set alone. Before believing the violation,
                 * recompute it from there, once per working set (see rr_active_set in sot_kernel_rr.cuh): the members'
                 * rows of N are lower triangular in the first q columns and R holds them signed and transposed, so
                 * R' z = S (bound - y_prev) by forward substitution, then y = y_prev + N z for every row. */
#pragma unroll 1
                for (int k = lane; k < q; k += 32) {
                    const int a = c.act[k], row = a & 0xffff;
                    const Real limk = (a & 0x10000) ? c.yhi[row] : c.ylo[row];
                    c.r[k] = ((a & 0x10000) ? -Real(1.0) : Real(1.0)) * (limk - c.yprev[row]);
                }
                __syncwarp();
#pragma unroll 1
                for (int k = 0; k < q; ++k) {
                    Real part = Real(0.0);
#pragma unroll 1
                    for (int j = lane; j < k; j += 32) part = fma(c.R[j * LD + k], c.dm[mo + j], part);
                    const Real sum = warp_sum<Real>(part);
                    __syncwarp();
                    if (lane == 0) c.dm[mo + k] = (c.r[k] - sum) * c.rinv[k];
                    __syncwarp();
                }
#pragma unroll 1
                for (int i = lane; i < NR; i += 32) {
                    const Real* row = c.T + i * LD + mo;
                    Real acc = Real(0.0);
#pragma unroll 1
                    for (int j = 0; j < q; ++j) acc = fma(row[j], c.dm[mo + j], acc);
                    c.y[i] = c.yprev[i] + acc;
                }
                refreshed = true;
                __syncwarp();
                break;      /* price again */
            }
            if (!(t < INFINITY)) {
                /* neither a primal nor a dual step exists. At a degenerate vertex (more rows
                 * tight than free directions, the normal case once earlier levels have
                 * saturated bounds) a row can be "violated" by accumulated rounding only:
                 * park it until the iterate moves again. Anything larger is a genuinely
                 * inconsistent constraint set. */
                if (-s <= Tr<Real>::degenerate_tol * fmax(Real(1.0), fabs(lim))) {
                    if (lane == 0) c.where[ip] = -2;
                    parked = true;
                    __syncwarp();
                    break;
                }
                out.status = OSB_INFEASIBLE; return out;
            }
            __syncwarp();   /* every lane has read y[ip] before its owner moves it */
            if (indep) { /* primal + dual step */
#pragma unroll 1
                for (int i = lane; i < NR; i += 32) c.y[i] = fma(t, c.dy[i], c.y[i]);
            }
#pragma unroll 1
            for (int k = lane; k < q; k += 32) c.u[k] = fma(-t, c.r[k], c.u[k]);
            up += t;
            PROF_T(g5);
            if (indep && t2 <= t1) {
                /* ---- full step: append ip. Householder maps d2 onto rho e_q ---- */
                const Real dq = c.d[mo + q];
                const Real rs = rsqrt(zz), sq = zz * rs;         /* 1/|d2|, |d2| */
                const Real rho = (dq > Real(0.0)) ? -sq : sq;
                const Real rho_inv = (dq > Real(0.0)) ? -rs : rs;
                const Real beta = Tr<Real>::rcp(rho * (rho - dq));
                __syncwarp();                                       /* every lane has read d[q], y */
                if (lane == 0) c.dm[mo + q] = dq - rho;             /* v = d2 - rho e1 */
                __syncwarp();
                {
                    const typename Tr<Real>::V2* dv = reinterpret_cast<const typename Tr<Real>::V2*>(c.dm + j0);
                    const int n2 = (wE - j0) >> 1;
#pragma unroll 1
                    for (int i = lane; i < NR; i += 32) {
                        Real* rowq = c.T + i * LD;
                        const Real w = beta * (c.dy[i] - rho * rowq[mo + q]); /* (N2 v)_i */
                        axpy_pairs<Real>(reinterpret_cast<typename Tr<Real>::V2*>(rowq + j0), dv, w, n2);
                    }
                }
#pragma unroll 1
                for (int k = lane; k < q; k += 32) c.R[k * LD + q] = c.d[mo + k];
                if (lane == 0) {
                    c.R[q * LD + q] = rho; c.rinv[q] = rho_inv;
                    c.act[q] = bidx; c.where[ip] = q; c.u[q] = up;
                }
                c.q = q + 1;
                refreshed = false;
                if (parked) { /* the iterate moved: parked rows are priced again */
#pragma unroll 1
                    for (int i = lane; i < NR; i += 32) if (c.where[i] == -2) c.where[i] = -1;
                    parked = false;
                }
                __syncwarp();
                PROF_T(g6);
                break;
            }
            __syncwarp();
            /* ---- partial (or pure dual) step: drop working-set member lk ---- */
            {
                const int k0 = lk;
                if (lane == 0) c.where[c.act[k0] & 0xffff] = -1;
                __syncwarp();
                /* delete column k0 of R: every row shifts its tail left by one (lane per row) */
#pragma unroll 1
                for (int i = lane; i < q; i += 32) {
                    Real* row = c.R + i * LD;
                    const int from = (i - 1 > k0) ? i - 1 : k0;
#pragma unroll 1
                    for (int j = from; j < q - 1; ++j) row[j] = row[j + 1];
                }
                Real un = 0; int an = 0;
#pragma unroll 1
                for (int base = k0; base < q - 1; base += 32) {
                    const int k = base + lane;
                    if (k < q - 1) { un = c.u[k + 1]; an = c.act[k + 1]; }
                    __syncwarp();
                    if (k < q - 1) { c.u[k] = un; c.act[k] = an; c.where[an & 0xffff] = k; }
                    __syncwarp();
                }
                __syncwarp();
                /* Givens rotations restore the triangle; same rotations on columns of N */
#pragma unroll 1
                for (int j = k0; j < q - 1; ++j) {
                    Real* Rj = c.R + j * LD;
                    Real* Rj1 = Rj + LD;
                    const Real a = Rj[j], b = Rj1[j];
                    if (b != Real(0.0)) {
                        const Real hinv = rsqrt(fma(a, a, b * b)), cs = a * hinv, sn = b * hinv;
                        __syncwarp();
#pragma unroll 1
                        for (int cc = j + lane; cc < q - 1; cc += 32) {
                            const Real ra = Rj[cc], rb = Rj1[cc];
                            Rj[cc] = cs * ra + sn * rb;
                            Rj1[cc] = cs * rb - sn * ra;
                        }
#pragma unroll 1
                        for (int i = lane; i < NR; i += 32) {
                            Real* row = c.T + i * LD + mo + j;
                            const Real ja = row[0], jb = row[1];
                            row[0] = cs * ja + sn * jb;
                            row[1] = cs * jb - sn * ja;
                        }
                        __syncwarp();
                    }
                    if (lane == 0) c.rinv[j] = Tr<Real>::rcp(Rj[j]);
                }
                c.q = q - 1;
                refreshed = false;
                __syncwarp();
            }
        }
    }
    return out;
}

} // namespace

template <typename Real>
__device__ __forceinline__ void sot_solve_impl(const sot_params& p)
{
    OSB_SMEM_RAW(ws_raw);
    Real* ws = reinterpret_cast<Real*>(ws_raw);
    const int lane = threadIdx.x;
    const int inst = blockIdx.x;
    if (inst >= p.batch) return;
    const osb_stack_desc& D = p.d;
    const osb_layout& L = p.L;
    const sot_workspace& W = p.w;

    PROF_T(t_start);
#ifdef OSB_PROFILE
    long long prof[OSB_MAX_LEVELS][7] = {};
#endif
    Ctx<Real> c;
    c.p = &p; c.lane = lane;
    c.blob = ws + W.off_blob; c.ZB = ws + W.off_ZB; c.T = ws + W.off_T; c.R = ws + W.off_R;
    c.X = ws + W.off_X;
    c.yprev = ws + W.off_yprev; c.y = ws + W.off_y; c.ylo = ws + W.off_ylo; c.yhi = ws + W.off_yhi;
    c.dy = ws + W.off_dy; c.d = ws + W.off_d; c.r = ws + W.off_r; c.u = ws + W.off_u;
    c.rinv = ws + W.off_rinv; c.gz = ws + W.off_gz; c.res = ws + W.off_res;
    c.dm = ws + W.off_dm; c.rowsq = ws + W.off_rowsq;
    c.rn2 = c.res + W.AUX + (W.AUX & 1);
    c.act = (int32_t*)(ws + W.off_act); c.where = (int32_t*)(ws + W.off_where);
    c.piv = (int32_t*)(ws + W.off_piv);
    const int LD = W.LD;
    c.LD = LD;
    /* elements between instance blobs: 16-byte multiples for the bulk copy */
    const int stride_e = (sizeof(Real) == 8) ? L.stride : ((L.stride + 3) & ~3);
    Real* const out_x = static_cast<Real*>(p.x);
    Real* const out_xl = static_cast<Real*>(p.x_levels);
    const int n = D.n, cin = D.c_in;
    c.n = n; c.cin = cin; c.NR0 = n + cin; c.z0 = 0; c.nz = n; c.q = 0; c.mo = 0;

    /* ---- stage the instance blob: one TMA bulk copy ---- */
    blob_stage(c.blob, static_cast<const Real*>(p.blobs) + (size_t)inst * stride_e,
               (uint32_t)stride_e * (uint32_t)sizeof(Real), ws + W.off_bar, lane);
    /* overlap: initialise the basis while the copy is in flight (lane per row) */
#pragma unroll 1
    for (int i = lane; i < n + cin; i += 32) {
        Real* row = c.ZB + i * LD;
#pragma unroll 1
        for (int j = 0; j < LD; ++j) row[j] = (i == j) ? Real(1.0) : Real(0.0); /* Cin rows are filled below */
    }
#pragma unroll 1
    for (int i = lane; i < W.NRmax; i += 32) {
        c.yprev[i] = Real(0.0);
        for (int j = n; j < LD; ++j) c.T[i * LD + j] = Real(0.0); /* pad columns take part in 128-bit row walks */
    }
    PROF_T(t_init);
    blob_wait(ws + W.off_bar);
    PROF_T(t_load);
    const Real* blob = c.blob;

    /* ---- bounds: l = max l_s, u = min u_s, clipped to +-1e20 ---- */
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        Real lo = -Real(OSB_INFTY), hi = Real(OSB_INFTY);
        for (int s = 0; s < D.n_bound_sets; ++s) {
            lo = fmax(lo, blob[L.l[s] + i]);
            hi = fmin(hi, blob[L.u[s] + i]);
        }
        c.ylo[i] = clip_inf<Real>(lo); c.yhi[i] = clip_inf<Real>(hi);
    }
#pragma unroll 1
    for (int r = lane; r < cin; r += 32) {
        c.ylo[n + r] = clip_inf<Real>(blob[L.lA + r]);
        c.yhi[n + r] = clip_inf<Real>(blob[L.uA + r]);
    }
#pragma unroll 1
    for (int r = 0; r < cin; ++r)
#pragma unroll 1
        for (int j = lane; j < n; j += 32) c.ZB[(n + r) * LD + j] = blob[L.Cin + r * n + j];
    __syncwarp();

    int status = OSB_OK, total_iters = 0;
    {   /* a NaN anywhere in the instance poisons everything downstream: flag it and stop */
        bool bad = false;
#pragma unroll 1
        for (int i = lane; i < L.stride; i += 32) { const Real v = blob[i]; bad |= (v != v); }
        if (__any_sync(FULL, bad)) status = OSB_NAN_INPUT;
    }

    /* ---- global equalities Ceq x = deq: eliminated before any level ---- */
    if (status != OSB_OK) {
#pragma unroll 1
        for (int i = lane; i < n; i += 32) { out_x[(size_t)inst * n + i] = Real(0.0); peer_store_x<Real>(p, inst, n, i, Real(0.0)); }
        if (lane == 0) { p.status[inst] = status; peer_store_status(p, inst, status); if (p.iters) p.iters[inst] = 0; }
        return;
    }
    if (D.c_eq > 0) {
        const int ce = D.c_eq;
        Real* Mb = c.R; /* z0 = 0: already in phase with the basis rows */
#pragma unroll 1
        for (int r = 0; r < ce; ++r)
#pragma unroll 1
            for (int j = lane; j < n; j += 32) Mb[r * LD + j] = blob[L.Ceq + r * n + j];
        __syncwarp();
        const int rank = nullspace_sweep<Real>(c, Mb, ce);
        /* particular solution: forward substitution on the pivot rows, consistency check
         * on the dependent ones (all lanes redundantly, ce is small) */
        int npiv = 0;
#pragma unroll 1
        for (int r = 0; r < ce; ++r) {
            Real s = blob[L.deq + r];
            const int pc = c.piv[r];
            const int upto = (pc >= 0) ? pc : npiv;
            if (pc >= 0) ++npiv;
            for (int k = 0; k < upto; ++k) s -= Mb[r * LD + k] * c.gz[k];
            if (pc >= 0) { if (lane == 0) c.gz[pc] = s / Mb[r * LD + pc]; }
            else if (fabs(s) > Tr<Real>::eq_tol * fmax(Real(1.0), fabs(blob[L.deq + r]))) status = OSB_INFEASIBLE;
            __syncwarp();
        }
#pragma unroll 1
        for (int i = lane; i < c.NR0; i += 32) c.yprev[i] = dot_cc<Real>(c.ZB + i * LD, c.gz, rank);
        c.z0 = rank; c.nz = n - rank;
        __syncwarp();
    }

    int last_level = -1;
    for (int k = 0; k < D.n_levels; ++k) if (p.active_mask & (1u << k)) last_level = k;

#pragma unroll 1
    for (int k = 0; k < D.n_levels && status == OSB_OK; ++k) {
        if (!(p.active_mask & (1u << k))) continue;
        const Real eps = (Real)p.eps_level[k];
        const int kind = D.task_kind[k], wk = D.weight_kind[k];
        const int m = (kind == OSB_TASK_HESSIAN) ? 0 : D.m[k];
        const int cl = D.c_loc[k];
        c.NR = c.NR0 + cl;
        const int max_iter = p.max_iter > 0 ? p.max_iter : 10 * (c.NR + n) + 20;
        const int nz = c.nz;
        /* everything indexed by a basis column lives one word into its row/vector when z0 is odd,
         * so that it shares the 16-byte phase of the basis rows ZB[i][z0 + j] */
        const int mo = c.z0 & 1;
        c.mo = mo;
        Real* gzs = c.gz + mo;   /* gradient / p, by basis column */
        Real* Rs = c.R + mo;     /* reduced Hessian / its Cholesky factor, by basis column */

        if (nz == 0) { /* nothing left to decide: the solution is inherited */
            if (p.store_levels)
                for (int i = lane; i < n; i += 32)
                    out_xl[((size_t)inst * D.n_levels + k) * n + i] = c.yprev[i];
            continue;
        }
        /* ---- local rows: limits and value at x_prev ---- */
#pragma unroll 1
        for (int r = lane; r < cl; r += 32) {
            c.ylo[c.NR0 + r] = clip_inf<Real>(blob[L.lCl[k] + r]);
            c.yhi[c.NR0 + r] = clip_inf<Real>(blob[L.uCl[k] + r]);
        }
        if (cl > 0) {
            residual_rows<Real>(c, blob + L.Cl[k], n, nullptr, cl, c.yprev, c.yprev + c.NR0);
#pragma unroll 1
            for (int r = lane; r < cl; r += 32) c.yprev[c.NR0 + r] = -c.yprev[c.NR0 + r];
        }
        /*
         * Three ways to factor the reduced Hessian Hz = L L' of the level:
         *  TRIVIAL  A = I, W = I:  Hz = (1+eps) I, nothing to factor.
         *  LQ       dense A, diagonal W: rotate the basis FIRST with the Householder sweep of
         *           M = A Z  (M Q = [Lm 0]); in the rotated basis Hz = blkdiag(Lm'W Lm + eps I,
         *           eps I): only a rank x rank Cholesky, and the same sweep is the basis
         *           shrink the next level needs.
         *  DENSE    everything else (H given, dense W, weighted identity, user regularisation
         *           task): form Hz explicitly, full Cholesky, sweep afterwards.
         */
        PROF_T(t_l0);
        const int path = sot_level_path_of(&D, k);
        int ncol = nz;             /* leading columns that need the Cholesky factor           */
        Real tail_scale = Real(1.0);   /* remaining columns: L is tail_scale^-1 * I               */
        int rank_lq = 0;
        Real* Mrows = nullptr;   /* task rows in basis coordinates (for the sweep)          */

        if (path == SOT_PATH_TRIVIAL) {
            ncol = 0; tail_scale = Real(1.0) / sqrt(Real(1.0) + eps);
#pragma unroll 1
            for (int i = lane; i < n; i += 32)
                c.res[i] = (Real(1.0) + eps) * c.yprev[i] - blob[L.b[k] + i] + (D.has_c[k] ? blob[L.c[k] + i] : Real(0.0));
            __syncwarp();
#pragma unroll 1
            for (int j = lane; j < nz; j += 32) gzs[j] = dot_cs<Real>(c.res, c.ZB + c.z0 + j, LD, n);
        } else if (path == SOT_PATH_LQ) {
            const int moff = (m < n) ? m : n; /* rows [0,moff) of R hold G, task rows follow */
            Mrows = c.R + moff * LD + (c.z0 & 1);
            if (c.z0 == 0 && nz == n && D.c_eq == 0) { /* basis still the identity: M = A */
#pragma unroll 1
                for (int r = 0; r < m; ++r)
#pragma unroll 1
                    for (int j = lane; j < n; j += 32) Mrows[r * LD + j] = blob[L.A[k] + r * n + j];
            } else {
                project_rows<Real>(c, blob + L.A[k], n, m, Mrows);
            }
            residual_rows<Real>(c, blob + L.A[k], n, blob + L.b[k], m, c.yprev, c.res);
            __syncwarp();
            PROF_T(t_proj);
            rank_lq = nullspace_sweep<Real>(c, Mrows, m);
            PROF_T(t_sweep);
#ifdef OSB_PROFILE
            prof[k][0] = t_proj - t_l0; prof[k][1] = t_sweep - t_proj;
#endif
            ncol = rank_lq; tail_scale = Real(1.0) / sqrt(eps);
            const Real* wd = (wk == OSB_W_DIAG) ? blob + L.w[k] : nullptr;
#pragma unroll 1
            for (int a = lane; a < ncol; a += 32) {
                Real* row = Rs + a * LD;
                if (mo) row[-1] = Real(0.0);
                for (int b = 0; b <= a; ++b) row[b] = (a == b) ? eps : Real(0.0);
            }
            __syncwarp();
            gram_pairs<Real>(c, Rs, Mrows, Mrows, m, wd, ncol);
            /* gz = eps Z' x_prev (+ Z' c) - Lm' W res */
            if (D.has_c[k]) {
#pragma unroll 1
                for (int i = lane; i < n; i += 32) c.rn2[i] = eps * c.yprev[i] + blob[L.c[k] + i];
            } else {
#pragma unroll 1
                for (int i = lane; i < n; i += 32) c.rn2[i] = eps * c.yprev[i];
            }
#pragma unroll 1
            for (int r = lane; r < m; r += 32) c.res[r] *= (wd ? wd[r] : Real(1.0));
            __syncwarp();
#pragma unroll 1
            for (int j = lane; j < nz; j += 32) {
                Real s = dot_cs<Real>(c.rn2, c.ZB + c.z0 + j, LD, n);
                if (j < ncol) s -= dot_cs<Real>(c.res, Mrows + j, LD, m);
                gzs[j] = s;
            }
        } else {
            /* ---- DENSE: Hz (lower triangle, in R) and gz explicitly ---- */
#pragma unroll 1
            for (int a = lane; a < nz; a += 32) {
                Real* row = Rs + a * LD;
                if (mo) row[-1] = Real(0.0);
                for (int b = 0; b <= a; ++b) row[b] = (a == b) ? eps : Real(0.0);
            }
            if (D.has_c[k]) {
#pragma unroll 1
                for (int i = lane; i < n; i += 32) c.rn2[i] = eps * c.yprev[i] + blob[L.c[k] + i];
            } else {
#pragma unroll 1
                for (int i = lane; i < n; i += 32) c.rn2[i] = eps * c.yprev[i];
            }
            __syncwarp();
#pragma unroll 1
            for (int j = lane; j < nz; j += 32) gzs[j] = dot_cs<Real>(c.rn2, c.ZB + c.z0 + j, LD, n);
            __syncwarp();
            const Real* Zx = c.ZB + c.z0; /* x rows of the basis, row stride LD */
            if (kind == OSB_TASK_DENSE) {
                Mrows = c.X + (c.z0 & 1);
                project_rows<Real>(c, blob + L.A[k], n, m, Mrows);
                residual_rows<Real>(c, blob + L.A[k], n, blob + L.b[k], m, c.yprev, c.res);
                __syncwarp();
                if (wk == OSB_W_DENSE) { /* WM and W res */
                    Real* WM = c.X + m * LD + 2;
                    const Real* Wm = blob + L.w[k];
#pragma unroll 1
                    for (int j = lane; j < nz; j += 32)
                        for (int r = 0; r < m; ++r) WM[r * LD + j] = dot_cs<Real>(Wm + r * m, Mrows + j, LD, m);
#pragma unroll 1
                    for (int r = lane; r < m; r += 32) c.rn2[r] = dot_cc<Real>(Wm + r * m, c.res, m);
                    __syncwarp();
                    gram_pairs<Real>(c, Rs, Mrows, WM, m, nullptr, nz);
#pragma unroll 1
                    for (int j = lane; j < nz; j += 32) gzs[j] -= dot_cs<Real>(c.rn2, Mrows + j, LD, m);
                } else {
                    const Real* wd = (wk == OSB_W_DIAG) ? blob + L.w[k] : nullptr;
                    gram_pairs<Real>(c, Rs, Mrows, Mrows, m, wd, nz);
#pragma unroll 1
                    for (int r = lane; r < m; r += 32) c.res[r] *= (wd ? wd[r] : Real(1.0));
                    __syncwarp();
#pragma unroll 1
                    for (int j = lane; j < nz; j += 32) gzs[j] -= dot_cs<Real>(c.res, Mrows + j, LD, m);
                }
            } else if (kind == OSB_TASK_IDENTITY) {
                /* A = I: task rows in basis coordinates are the x rows of ZB themselves */
#pragma unroll 1
                for (int i = lane; i < n; i += 32) c.res[i] = blob[L.b[k] + i] - c.yprev[i];
                __syncwarp();
                if (wk == OSB_W_DENSE) {
                    Real* WM = c.X;
                    const Real* Wm = blob + L.w[k];
#pragma unroll 1
                    for (int j = lane; j < nz; j += 32)
                        for (int r = 0; r < n; ++r) WM[r * LD + j] = dot_cs<Real>(Wm + r * n, Zx + j, LD, n);
#pragma unroll 1
                    for (int r = lane; r < n; r += 32) c.rn2[r] = dot_cc<Real>(Wm + r * n, c.res, n);
                    __syncwarp();
                    gram_pairs<Real>(c, Rs, Zx, WM, n, nullptr, nz);
#pragma unroll 1
                    for (int j = lane; j < nz; j += 32) gzs[j] -= dot_cs<Real>(c.rn2, Zx + j, LD, n);
                } else {
                    const Real* wd = (wk == OSB_W_DIAG) ? blob + L.w[k] : nullptr;
                    gram_pairs<Real>(c, Rs, Zx, Zx, n, wd, nz);
#pragma unroll 1
                    for (int r = lane; r < n; r += 32) c.res[r] *= (wd ? wd[r] : Real(1.0));
                    __syncwarp();
#pragma unroll 1
                    for (int j = lane; j < nz; j += 32) gzs[j] -= dot_cs<Real>(c.res, Zx + j, LD, n);
                }
            } else { /* OSB_TASK_HESSIAN: Hz = Z' H Z, gz += Z'(H x_prev + g) */
                const Real* Hm = blob + L.A[k];
                Real* HZ = c.X;
                project_rows<Real>(c, Hm, n, n, HZ);
                residual_rows<Real>(c, Hm, n, nullptr, n, c.yprev, c.res); /* -H x_prev */
#pragma unroll 1
                for (int i = lane; i < n; i += 32) c.res[i] = blob[L.b[k] + i] - c.res[i];
                __syncwarp();
                gram_pairs<Real>(c, Rs, Zx, HZ, n, nullptr, nz);
#pragma unroll 1
                for (int j = lane; j < nz; j += 32) gzs[j] += dot_cs<Real>(c.res, Zx + j, LD, n);
            }
            __syncwarp();
            if (D.m_reg > 0) { /* user regularisation task: cost only, never optimality rows */
                const int mr = D.m_reg;
                Real* Mr = c.X + (W.XR - mr) * LD;
                project_rows<Real>(c, blob + L.Ar, n, mr, Mr);
                residual_rows<Real>(c, blob + L.Ar, n, blob + L.br, mr, c.yprev, c.rn2);
                __syncwarp();
                const Real* wd = (D.reg_weight_kind == OSB_W_DIAG) ? blob + L.wr : nullptr;
                gram_pairs<Real>(c, Rs, Mr, Mr, mr, wd, nz);
#pragma unroll 1
                for (int r = lane; r < mr; r += 32) c.rn2[r] *= (wd ? wd[r] : Real(1.0));
                __syncwarp();
#pragma unroll 1
                for (int j = lane; j < nz; j += 32) gzs[j] -= dot_cs<Real>(c.rn2, Mr + j, LD, mr);
            }
        }
        __syncwarp();
        PROF_T(t_cost);
        /* ---- Cholesky of the leading ncol x ncol block (left-looking, lane per row) ---- */
        bool spd = true;
#pragma unroll 1
        for (int col = 0; col < ncol; ++col) {
            const Real* Rc = c.R + col * LD; /* word 0 of row col; column b at word mo + b */
            /* lane per row i >= col: s = G[i][col] - L[i][:col] . L[col][:col]; the pivot travels
             * by shuffle from the lane that owns row `col`, so one barrier per column suffices */
            Real s0 = 0, s1 = 0;
            const int i0 = col + lane, i1 = i0 + 32;
            if (i0 < ncol) s0 = Rs[i0 * LD + col] - dot_prefix<Real>(c.R + i0 * LD, Rc, mo + col);
            if (i1 < ncol) s1 = Rs[i1 * LD + col] - dot_prefix<Real>(c.R + i1 * LD, Rc, mo + col);
            const Real piv = __shfl_sync(FULL, s0, 0);
            if (!(piv > Real(0.0))) { spd = false; break; }
            const Real inv = rsqrt(piv);
            if (i0 < ncol) Rs[i0 * LD + col] = s0 * inv; /* diagonal becomes sqrt(piv) */
            if (i1 < ncol) Rs[i1 * LD + col] = s1 * inv;
            if (lane == 0) c.rinv[col] = inv;             /* 1 / L[col][col] */
            __syncwarp();
        }
        if (!spd) { status = OSB_NOT_SPD | (k << 8); break; }
        /* ---- p = L^-1 gz ---- */
#pragma unroll 1
        for (int col = 0; col < ncol; ++col) {
            const Real pc = gzs[col] * c.rinv[col];
            __syncwarp();
            if (lane == 0) gzs[col] = pc;
#pragma unroll 1
            for (int i = col + 1 + lane; i < ncol; i += 32) gzs[i] -= Rs[i * LD + col] * pc;
            __syncwarp();
        }
#pragma unroll 1
        for (int j = ncol + lane; j < nz; j += 32) gzs[j] *= tail_scale;
        PROF_T(t_chol);
        /* ---- local rows in the (rotated) basis ---- */
        if (cl > 0) project_rows<Real>(c, blob + L.Cl[k], n, cl, c.T + c.NR0 * LD + mo);
        /* guard words of p, so that full-row pair walks may run over them */
        if (lane == 0) {
            if (mo) gzs[-1] = Real(0.0);
            if ((mo + nz) & 1) gzs[nz] = Real(0.0);
        }
        __syncwarp();
        /* ---- N = (E Z) L^-T (lane per row) and the unconstrained minimiser y = y_prev - N p ----
         * word w of a T row holds basis column w - mo. Basis rows: one scaled copy of the whole
         * row (tail columns are final after it), then the triangular solve of the leading ncol
         * columns reading the untouched originals from ZB; local rows (few) in place. */
        {
            const int wend = mo + nz;                 /* one past the last used word */
            const int n2 = (wend + 1) >> 1;           /* pairs of a full-row walk     */
#pragma unroll 1
            for (int i = lane; i < c.NR; i += 32) {
                Real* Tw = c.T + i * LD;            /* word 0 */
                Real* Ti = Tw + mo;                 /* column 0 */
                if (i < c.NR0) {
                    const Real* Zw = c.ZB + i * LD + c.z0 - mo; /* even word: same phase as Tw */
                    const typename Tr<Real>::V2* z2 = reinterpret_cast<const typename Tr<Real>::V2*>(Zw);
                    typename Tr<Real>::V2* t2 = reinterpret_cast<typename Tr<Real>::V2*>(Tw);
#pragma unroll 1
                    for (int j = 0; j < n2; j += 4) {
                        const int j1 = min(j + 1, n2 - 1), j2 = min(j + 2, n2 - 1), j3 = min(j + 3, n2 - 1);
                        typename Tr<Real>::V2 a0 = z2[j], a1 = z2[j1], a2 = z2[j2], a3 = z2[j3];
                        a0.x *= tail_scale; a0.y *= tail_scale; a1.x *= tail_scale; a1.y *= tail_scale;
                        a2.x *= tail_scale; a2.y *= tail_scale; a3.x *= tail_scale; a3.y *= tail_scale;
                        t2[j] = a0; t2[j1] = a1; t2[j2] = a2; t2[j3] = a3;
                    }
                    const Real* Zi = Zw + mo;
#pragma unroll 1
                    for (int col = 0; col < ncol; ++col)
                        Ti[col] = (Zi[col] - dot_prefix<Real>(Tw, c.R + col * LD, mo + col)) * c.rinv[col];
                } else {
                    if (mo) Tw[0] = Real(0.0);
#pragma unroll 1
                    for (int col = 0; col < ncol; ++col)
                        Ti[col] = (Ti[col] - dot_prefix<Real>(Tw, c.R + col * LD, mo + col)) * c.rinv[col];
#pragma unroll 1
                    for (int col = ncol; col < nz; ++col) Ti[col] *= tail_scale;
                }
                if (mo) Tw[0] = Real(0.0);                  /* guard / pad words of the pair walks */
                if (wend & 1) Tw[wend] = Real(0.0);
                /* y = y_prev - N p and |row|^2 in one walk */
                const typename Tr<Real>::V2* t2 = reinterpret_cast<const typename Tr<Real>::V2*>(Tw);
                const typename Tr<Real>::V2* g2 = reinterpret_cast<const typename Tr<Real>::V2*>(c.gz);
                Real acc0 = 0, acc1 = 0, sq0 = 0, sq1 = 0;
#pragma unroll 1
                for (int j = 0; j < n2; j += 4) {
                    const int j1 = min(j + 1, n2 - 1), j2 = min(j + 2, n2 - 1), j3 = min(j + 3, n2 - 1);
                    const typename Tr<Real>::V2 a0 = t2[j], a1 = t2[j1], a2 = t2[j2], a3 = t2[j3];
                    const typename Tr<Real>::V2 b0 = g2[j], b1 = g2[j1], b2 = g2[j2], b3 = g2[j3];
                    acc0 = fma(a0.x, b0.x, acc0); acc1 = fma(a0.y, b0.y, acc1);
                    sq0 = fma(a0.x, a0.x, sq0); sq1 = fma(a0.y, a0.y, sq1);
                    if (j + 1 < n2) { acc0 = fma(a1.x, b1.x, acc0); acc1 = fma(a1.y, b1.y, acc1);
                                      sq0 = fma(a1.x, a1.x, sq0); sq1 = fma(a1.y, a1.y, sq1); }
                    if (j + 2 < n2) { acc0 = fma(a2.x, b2.x, acc0); acc1 = fma(a2.y, b2.y, acc1);
                                      sq0 = fma(a2.x, a2.x, sq0); sq1 = fma(a2.y, a2.y, sq1); }
                    if (j + 3 < n2) { acc0 = fma(a3.x, b3.x, acc0); acc1 = fma(a3.y, b3.y, acc1);
                                      sq0 = fma(a3.x, a3.x, sq0); sq1 = fma(a3.y, a3.y, sq1); }
                }
                c.y[i] = c.yprev[i] - (acc0 + acc1);
                c.rowsq[i] = sq0 + sq1; /* |row|^2 of N: invariant under the active-set column rotations */
            }
        }
        __syncwarp();
        /* ---- inequality rows and bounds ---- */
        PROF_T(t_nb);
        LevelOut lo = dual_active_set<Real>(c, max_iter);
        PROF_T(t_gi);
#ifdef OSB_PROFILE
        prof[k][2] = t_cost - t_l0; prof[k][3] = t_chol - t_cost; prof[k][4] = t_nb - t_chol;
        prof[k][5] = t_gi - t_nb; prof[k][6] = lo.iters;
#endif
        total_iters += lo.iters;
        if (lo.status != OSB_OK) { status = lo.status | (k << 8); break; }
        /* ---- level solution; active bounds are returned exactly on the bound ---- */
#pragma unroll 1
        for (int i = lane; i < c.NR0; i += 32) {
            Real v = c.y[i];
            if (i < n && c.where[i] >= 0) v = (c.act[c.where[i]] & 0x10000) ? c.yhi[i] : c.ylo[i];
            c.yprev[i] = v;
            /* later levels keep this point feasible (see sot_kernel_rr.cuh): their box contains it */
            c.ylo[i] = fmin(c.ylo[i], v); c.yhi[i] = fmax(c.yhi[i], v);
            if (p.store_levels && i < n) out_xl[((size_t)inst * D.n_levels + k) * n + i] = v;
        }
        __syncwarp();
        /* ---- shrink the basis by the row space of this level's task ---- */
        if (path == SOT_PATH_LQ) { c.z0 += rank_lq; c.nz -= rank_lq; }
        else if (k != last_level) {
            if (kind == OSB_TASK_IDENTITY) { c.z0 += c.nz; c.nz = 0; }
            else if (kind == OSB_TASK_DENSE) {
                const int rank = nullspace_sweep<Real>(c, Mrows, m);
                c.z0 += rank; c.nz -= rank;
            }
        }
    }
#ifdef OSB_PROFILE
    {
        const long long t_end = clock64();
        if (threadIdx.x == 0 && blockIdx.x == OSB_PROFILE) {
            printf("total %lld init %lld loadwait %lld\n", t_end - t_start, t_init - t_start, t_load - t_init);
            for (int k = 0; k < D.n_levels; ++k)
                printf("L%d project %lld sweep %lld | cost(all) %lld chol+fwd %lld Nbuild %lld GI %lld (%lld iters)\n", k,
                       prof[k][0], prof[k][1], prof[k][2], prof[k][3], prof[k][4], prof[k][5], prof[k][6]);
        }
    }
#endif
    /* ---- result ---- */
#pragma unroll 1
    for (int i = lane; i < n; i += 32) { out_x[(size_t)inst * n + i] = c.yprev[i]; peer_store_x<Real>(p, inst, n, i, c.yprev[i]); }
    if (lane == 0) {
        p.status[inst] = status;
        peer_store_status(p, inst, status);
        if (p.iters) p.iters[inst] = total_iters;
    }
}


#ifndef OSB_EMU
extern "C" __global__ void __launch_bounds__(32) sot_solve_kernel(const __grid_constant__ sot_params p)
{
    sot_solve_impl<double>(p);
}
extern "C" __global__ void __launch_bounds__(32) sot_solve_kernel_f32(const __grid_constant__ sot_params p)
{
    sot_solve_impl<float>(p);
}

/* ---- Aggregated / AutoStack::update() on device: gather child blocks into blobs ------ */
extern "C" __global__ void __launch_bounds__(256)
sot_gather_kernel(const __grid_constant__ gather_params g)
{
    /* One CTA walks whole instances (grid-stride), all sources of an instance back to back: consecutive threads
     * read consecutive source doubles and write consecutive blob doubles. A source whose rows are contiguous and
     * taken in order is one flat copy (no division per element). (First version: one 256-thread CTA per
     * (instance, source) pair of 87..348 elements with a division per element: 1.85 TB/s, 29 % of the HBM copy
     * peak at B = 65536; see profiles/r02_aux_kernels*.jsonl.) */
    for (int inst = blockIdx.x; inst < g.batch; inst += gridDim.x) {
        double* blob = g.blobs + (size_t)inst * g.stride;
#pragma unroll 1
        for (int s = 0; s < g.n_srcs; ++s) {
            const osb_gather_src& S = g.src[s];
            const double* src = S.src + (size_t)inst * S.src_stride;
            double* dst = blob + S.dst_off;
            const int total = S.rows * S.cols;
            const double scale = S.scale;
            if (!S.row_index && S.src_ld == S.cols) {
                for (int e = threadIdx.x; e < total; e += blockDim.x) dst[e] = scale * src[e];
            } else {
                for (int e = threadIdx.x; e < total; e += blockDim.x) {
                    const int r = e / S.cols, cidx = e - r * S.cols;
                    const int sr = S.row_index ? S.row_index[r] : r;
                    dst[e] = scale * src[(size_t)sr * S.src_ld + cidx];
                }
            }
        }
    }
}

/* ---- joint-state producers: Postural error, JointLimits and VelocityLimits boxes (HBM-bound) ---- */
extern "C" __global__ void __launch_bounds__(256)
sot_posture_kernel(const __grid_constant__ posture_params g)
{
    /* one thread per (instance, joint): consecutive threads read consecutive joints of q and write
     * consecutive doubles of each destination field */
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)g.batch * g.n) return;
    const int inst = (int)(e / g.n), i = (int)(e - (int64_t)inst * g.n);
    const osb_posture_inputs& P = g.in;
    const double q = P.q[(size_t)inst * P.q_stride + i];
    double* blob = g.blobs + (size_t)inst * g.stride;
    if (g.off_b >= 0) {      /* Postural.cpp:97-100  b = v_ref + lambda (q_ref - q) */
        const double dq = P.q_ref[(size_t)inst * P.q_ref_stride + i] - q;
        const double v = P.v_ref ? P.v_ref[(size_t)inst * P.v_ref_stride + i] : 0.0;
        blob[g.off_b + i] = __dadd_rn(v, __dmul_rn(P.lambda, dq));   /* no FMA contraction: bit-identical to the host formula */
    }
    if (g.off_jl >= 0) {     /* JointLimits.cpp:37-59: scaled distance to the limits, 0 always inside */
        const double dq = q - (P.q_neutral ? P.q_neutral[i] : 0.0);
        blob[g.off_ju + i] = fmax((P.q_max[i] - dq) * P.bound_scaling, 0.0);
        blob[g.off_jl + i] = fmin((P.q_min[i] - dq) * P.bound_scaling, 0.0);
    }
    if (g.off_vl >= 0) {     /* VelocityLimits.cpp:72-89 */
        const double v = fabs(P.qdot_max[i]) * P.dT;
        blob[g.off_vl + i] = -1.0 * v;
        blob[g.off_vu + i] = 1.0 * v;
    }
}

/* ---- device-side barrier of the fused all-gather (osb_gather_barrier) ----------------------
 * Stream-ordered after this rank's solve kernel, so its peer stores have been issued; lane r
 * publishes `epoch` into word [rank] of rank r's flag block (system-scope fence first: the
 * rows written by the solve kernel are visible before the flag), then waits until rank r's
 * epoch has arrived in the own block. A peer that never arrives ends the wait after
 * timeout_cycles and is counted in *fail: the GPU is never left spinning. */
extern "C" __global__ void __launch_bounds__(32) sot_gather_barrier_kernel(const gather_barrier_params g)
{
    const int r = threadIdx.x;
    __threadfence_system();
    if (r < g.world && r != g.rank) {
        volatile uint32_t* theirs = g.flags[r] + g.rank;
        *theirs = g.epoch;
    }
    __threadfence_system();
    bool late = false;
    if (r < g.world && r != g.rank) {
        volatile uint32_t* mine = g.flags[g.rank] + r;
        const long long t0 = clock64();
        while ((int32_t)(*mine - g.epoch) < 0) {
            if (clock64() - t0 > g.timeout_cycles) { late = true; break; }
            __nanosleep(100);
        }
    }
    __threadfence_system();
    if (__any_sync(0xffffffffu, late) && r == 0) atomicAdd(g.fail, 1);
}

/* ---- FP64 FMA peak probe (roofline denominator) -------------------------------------- */
extern "C" __global__ void __launch_bounds__(256) sot_dfma_probe_kernel(double* out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, b = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
/* ---- FP32 FMA peak probe (roofline denominator of the fp32 engines) ------------------------- */
extern "C" __global__ void __launch_bounds__(256) sot_ffma_probe_kernel(float* out, int iters)
{
    float a0 = threadIdx.x * 1e-6f, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
          a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0001f, b = 1e-4f;
    for (int i = 0; i < iters; ++i) {
        a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
        a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
#endif /* !OSB_EMU */

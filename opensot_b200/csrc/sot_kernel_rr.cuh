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
 * What is different from sot_kernel.cuh is where the data lives and how the code is shaped:
 *  - a lane keeps its row (of the task block M during the LQ sweep, of the basis ZB while it
 *    takes the reflectors, of N = (E Z) F^-T during the active-set loop) in registers as NP
 *    pairs; an inner loop is one fully unrolled walk over those pairs against a broadcast
 *    vector (one LDS.128 per pair, no address arithmetic, independent FMAs); the staging
 *    copy T of the generic kernel is gone from shared memory (more resident warps);
 *  - column indices are ABSOLUTE (0..LD): the live basis is columns [z0, n) of ZB, vectors
 *    are zero outside the range they act on, so no walk needs a mask, a peel or a dynamic
 *    register index;
 *  - the code is kept SMALL on purpose. Measured on B200 (profiles/): with one warp per CTA
 *    and ~10 unsynchronised warps per SM the instruction stream of a 7 500-instruction build
 *    came from L2 (L0 holds ~380 instructions, L1.5 ~2 000): 27 % of all stall samples were
 *    "no instruction". Hence one sweep site serving equalities and levels, one reflect body,
 *    one left-looking pass for Cholesky / forward solve / leading columns of N, rolled scalar
 *    loops around the unrolled walks. A working-set removal is a plane-rotation update of the
 *    inverse factor and of the register rows (rr_active_set): re-factorising instead, tried
 *    first to save code, made stacks with many removals the stragglers of a launch.
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
/* development aid: cycle stamps of one warp (block OSB_RR_PROFILE prints them) */
#define RPROF_T(var) const long long var = clock64()
#define RPROF_ADD(acc, a, b) (acc) += (b) - (a)
#define RPROF_MARK(i) do { const long long now__ = clock64(); prof[i] += now__ - prof_last; prof_last = now__; } while (0)
__device__ long long rr_prof_sw[2];
__device__ long long rr_prof_gi[8];
#else
#define RPROF_MARK(i)
#define RPROF_T(var)
#define RPROF_ADD(acc, a, b)
#endif

namespace {

/* keeps the compiler from hoisting the loads of a later chunk of a walk over this point: bounds
 * the registers a fully unrolled walk holds in flight (spills cost an L2 round trip here: the
 * L1 is carved down to almost nothing by the shared-memory configuration) */
#ifdef OSB_EMU
#define RR_FENCE() do { } while (0)
#else
#define RR_FENCE() asm volatile("" ::: "memory")
#endif

/*
 * Everything one warp needs besides its register rows. Few base pointers on purpose (every
 * pointer is a live register for the whole kernel): the column-indexed vectors sit LD apart
 * behind `vec`, the row-indexed ones RS = 32 RPL apart behind `rowv`, both strides known at
 * compile time.
 */
template <typename Real>
struct RC {
    const sot_params* p;
    int lane;
    Real *blob, *ZB, *M, *Lc;
    Real *vec, *rowv;
    int32_t *act, *piv;
    int LDL;            /* row stride of the compact factor Lc */
    int n, NR0, NR;
    int z0;             /* live basis = columns [z0, n) */
};
/* column-indexed vectors (LD entries each) */
#define RR_D(c)      ((c).vec)               /* candidate row of N, signed                         */
#define RR_DM(c)     ((c).vec + LD)          /* ... masked to the free columns                     */
#define RR_VM(c)     ((c).vec + 2 * LD)      /* Householder vector of an append; Z' operand before */
#define RR_PV(c)     ((c).vec + 3 * LD)      /* gradient / p = F^-1 g_z                            */
#define RR_CSCALE(c) ((c).vec + 4 * LD)      /* column scales of N                                 */
#define RR_U(c)      ((c).vec + 5 * LD)      /* multipliers of the working set                     */
#define RR_OH(c)     ((c).vec + 6 * LD)      /* unit vector of the first free column               */
#define RR_RINV(c)   ((c).vec + 7 * LD)      /* 1 / F[c][c]                                        */
#define RR_RES(c)    RR_D(c)                 /* task residual (dead before the active-set loop)    */
#define RR_BETA(c)   RR_DM(c)                /* reflector scales (dead before the active-set loop) */
/* row-indexed vectors (RS entries each) */
#define RR_YPREV(c)  ((c).rowv)
#define RR_LO(c)     ((c).rowv + RS)
#define RR_HI(c)     ((c).rowv + 2 * RS)
#define RR_RQ(c)     ((c).rowv + 3 * RS)
#define RR_RP(c)     ((c).M)                 /* packed inverse working-set factor, over M          */

template <typename Real, int NP>
__device__ __forceinline__ void rr_load(typename Tr<Real>::V2 (&t)[NP], const Real* __restrict__ row)
{
    const typename Tr<Real>::V2* r2 = reinterpret_cast<const typename Tr<Real>::V2*>(row);
#pragma unroll
    for (int k = 0; k < NP; ++k) t[k] = r2[k];
}
/* a lane without a row holds zeros (it never publishes or stores; reading another lane's row instead
 * would be a read the owner's later store races with) */
template <typename Real, int NP>
__device__ __forceinline__ void rr_load_if(typename Tr<Real>::V2 (&t)[NP], const Real* __restrict__ row, bool has)
{
    if (has) rr_load<Real, NP>(t, row);
    else {
#pragma unroll
        for (int k = 0; k < NP; ++k) { t[k].x = Real(0.0); t[k].y = Real(0.0); }
    }
}
template <typename Real, int NP>
__device__ __forceinline__ void rr_store(Real* __restrict__ row, const typename Tr<Real>::V2 (&t)[NP])
{
    typename Tr<Real>::V2* r2 = reinterpret_cast<typename Tr<Real>::V2*>(row);
#pragma unroll
    for (int k = 0; k < NP; ++k) r2[k] = t[k];
}
/* static select of slot s out of a small per-slot register array */
template <typename T, int RPL>
__device__ __forceinline__ T rr_pick(const T (&a)[RPL], int s)
{
    T v = a[0];
#pragma unroll
    for (int k = 1; k < RPL; ++k) if (s >= k) v = a[k];   /* (no equality test: see rr_publish) */
    return v;
}
/* the owner lane of `row` writes its register row to dst (the other lanes do nothing) */
template <typename Real, int NP, int RPL>
__device__ __forceinline__ void rr_publish(const typename Tr<Real>::V2 (&t)[RPL][NP], int lane, int row, Real* dst)
{
    if (lane == (row & 31)) {
        /* range tests, not `(row >> 5) == s`: with an equality the compiler rewrites t[s] as
         * t[row >> 5], a dynamic index that pushes the register rows into local memory */
#pragma unroll
        for (int s = 0; s < RPL; ++s)
            if (row >= 32 * s && row < 32 * (s + 1)) rr_store<Real, NP>(dst, t[s]);
    }
}
/*
 * The one reflect body: every row of the lane (all RPL slots) takes H = I - beta u u'.
 * Two walks over the broadcast vector u: w = beta (u . x), then x -= w u.
 * `dead`: u is zero on columns [0, dead) (a reflector of a sweep starts at its pivot column): when that covers the first
 * KS pairs, both walks enter behind them (one warp-uniform branch each; the skipped terms are exact zeros, the result is
 * bit-identical). Used by the CTA-per-stack kernel in fp64 only, where it was measured to pay (sot_kernel_mw.cuh); the
 * register-row kernel passes 0 and compiles to the straight walks: there the same skip cost 1.5-3.4 %.
 */
#ifndef RR_KS_REFL
#define RR_KS_REFL 4
#endif
#ifndef RR_KS_AS
#define RR_KS_AS 7      /* the same for the walks of the active-set loop (mw_active_set): the last level of a 3 / 12 / n stack has z0 = 15 */
#endif
template <typename Real, int NP, int RPL>
__device__ __forceinline__ void rr_reflect(typename Tr<Real>::V2 (&t)[RPL][NP], const Real* __restrict__ u, Real beta, int dead = 0)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int KS = (sizeof(Real) == 8 && NP > 2 * RR_KS_REFL) ? RR_KS_REFL : 0;
    const V2* u2 = reinterpret_cast<const V2*>(u);
    const bool full = (KS == 0) || dead < 2 * KS;
    Real dx[RPL][2], dy[RPL][2];     /* four accumulation chains per row */
#pragma unroll
    for (int s = 0; s < RPL; ++s) { dx[s][0] = dx[s][1] = Real(0.0); dy[s][0] = dy[s][1] = Real(0.0); }
#define RR_REFLECT_DOT(k)                                                                              \
    {                                                                                                  \
        if (RPL > 1 && (k) % (sizeof(Real) == 8 ? 2 : 4) == 0 && (k) > 0) RR_FENCE();                    \
        const V2 uv = u2[k];                                                                           \
        _Pragma("unroll") for (int s = 0; s < RPL; ++s) {                                              \
            dx[s][(k) & 1] = fma(uv.x, t[s][k].x, dx[s][(k) & 1]);                                     \
            dy[s][(k) & 1] = fma(uv.y, t[s][k].y, dy[s][(k) & 1]);                                     \
        }                                                                                              \
    }
    if (full) {
#pragma unroll
        for (int k = 0; k < KS; ++k) RR_REFLECT_DOT(k)
    }
#pragma unroll
    for (int k = KS; k < NP; ++k) RR_REFLECT_DOT(k)
#undef RR_REFLECT_DOT
    Real w[RPL];
#pragma unroll
    for (int s = 0; s < RPL; ++s) w[s] = beta * ((dx[s][0] + dx[s][1]) + (dy[s][0] + dy[s][1]));
#define RR_REFLECT_UPD(k)                                                                              \
    {                                                                                                  \
        if (RPL > 1 && (k) % (sizeof(Real) == 8 ? 2 : 4) == 0 && (k) > 0) RR_FENCE();                    \
        const V2 uv = u2[k];                                                                           \
        _Pragma("unroll") for (int s = 0; s < RPL; ++s) {                                              \
            t[s][k].x = fma(-w[s], uv.x, t[s][k].x); t[s][k].y = fma(-w[s], uv.y, t[s][k].y);          \
        }                                                                                              \
    }
    if (full) {
#pragma unroll
        for (int k = 0; k < KS; ++k) RR_REFLECT_UPD(k)
    }
#pragma unroll
    for (int k = KS; k < NP; ++k) RR_REFLECT_UPD(k)
#undef RR_REFLECT_UPD
}

/* dot of two shared-memory runs of dynamic length (short: the leading block of a level) */
template <typename Real>
__device__ __forceinline__ Real rr_dot_dyn(const Real* __restrict__ a, const Real* __restrict__ b, int cnt)
{
    Real s0 = 0, s1 = 0;
    int j = 0;
#pragma unroll 1
    for (; j + 1 < cnt; j += 2) { s0 = fma(a[j], b[j], s0); s1 = fma(a[j + 1], b[j + 1], s1); }
    if (j < cnt) s0 = fma(a[j], b[j], s0);
    return s0 + s1;
}

/*
 * dst rows (row r at dst + r*LD, absolute columns, zero outside the live basis)
 *   <-  A (m x n row-major, row stride lda) times the basis, or A itself with the row stride
 * changed to LD while the basis is still the identity.  Lane per basis column, four rows of
 * A per walk down the column.
 */
template <typename Real, int LD>
__device__ __forceinline__ void rr_project(RC<Real>& c, const Real* __restrict__ A, int lda, int m, Real* dst, bool identity)
{
    const int n = c.n;
#pragma unroll 1
    for (int j = c.lane; j < LD; j += 32) {
        const bool live = (j >= c.z0) && (j < n);
        const Real* zcol = c.ZB + j;
#pragma unroll 1
        for (int r = 0; r < m; r += 4) {
            const int r1 = min(r + 1, m - 1), r2 = min(r + 2, m - 1), r3 = min(r + 3, m - 1);
            const Real *a0 = A + r * lda, *a1 = A + r1 * lda, *a2 = A + r2 * lda, *a3 = A + r3 * lda;
            Real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            if (identity) {
                if (live) { s0 = a0[j]; s1 = a1[j]; s2 = a2[j]; s3 = a3[j]; }
            } else if (live) {
                int i = 0;
#pragma unroll 1
                for (; i + 3 < n; i += 4) {     /* 20 loads, then 16 independent multiply-adds */
                    const Real* z = zcol + i * LD;
                    const Real z0_ = z[0], z1_ = z[LD], z2_ = z[2 * LD], z3_ = z[3 * LD];
                    const Real p0 = a0[i], p1 = a0[i + 1], p2 = a0[i + 2], p3 = a0[i + 3];
                    const Real q0 = a1[i], q1 = a1[i + 1], q2 = a1[i + 2], q3 = a1[i + 3];
                    const Real u0 = a2[i], u1 = a2[i + 1], u2 = a2[i + 2], u3 = a2[i + 3];
                    const Real v0 = a3[i], v1 = a3[i + 1], v2 = a3[i + 2], v3 = a3[i + 3];
                    s0 = fma(p0, z0_, s0); s1 = fma(q0, z0_, s1); s2 = fma(u0, z0_, s2); s3 = fma(v0, z0_, s3);
                    s0 = fma(p1, z1_, s0); s1 = fma(q1, z1_, s1); s2 = fma(u1, z1_, s2); s3 = fma(v1, z1_, s3);
                    s0 = fma(p2, z2_, s0); s1 = fma(q2, z2_, s1); s2 = fma(u2, z2_, s2); s3 = fma(v2, z2_, s3);
                    s0 = fma(p3, z3_, s0); s1 = fma(q3, z3_, s1); s2 = fma(u3, z3_, s2); s3 = fma(v3, z3_, s3);
                }
#pragma unroll 1
                for (; i < n; ++i) {
                    const Real zv = zcol[i * LD];
                    s0 = fma(a0[i], zv, s0); s1 = fma(a1[i], zv, s1);
                    s2 = fma(a2[i], zv, s2); s3 = fma(a3[i], zv, s3);
                }
            }
            dst[r * LD + j] = s0; dst[r1 * LD + j] = s1; dst[r2 * LD + j] = s2; dst[r3 * LD + j] = s3;
        }
    }
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
 * The same product M = A Z without the basis: Z (all n columns) is the product of the Householder
 * reflectors of the levels swept so far, so the rows of A (lane per row, registers) simply take
 * those reflectors in the order they were made -- O(m n rank) instead of O(m n n), and 30 loads +
 * 60 multiply-adds per reflector instead of a load per multiply-add. The reflectors were saved
 * by the level loop, scaled to I - u u', over the dead A block of their level (`ranks` holds the
 * count per level, eight bits each). Columns left of z0 (components along the earlier levels'
 * row spaces) are zeroed on the way to c.M.
 */
template <typename Real, int NP, int RPL>
__device__ __forceinline__ void rr_project_refl(RC<Real>& c, typename Tr<Real>::V2 (&t)[RPL][NP], const Real* __restrict__ A,
                                                int m, int level, unsigned long long ranks, const int32_t* offA)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP;
    const int lane = c.lane, n = c.n, z0 = c.z0;
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        const Real* src = A + ((row < m) ? row : 0) * n;
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            t[s][k].x = (2 * k < n) ? src[2 * k] : Real(0.0);
            t[s][k].y = (2 * k + 1 < n) ? src[2 * k + 1] : Real(0.0);
        }
    }
    Real* W = RR_VM(c);
#pragma unroll 1
    for (int j = 0; j < level; ++j) {
        const int cnt = (int)((ranks >> (8 * j)) & 0xffu);
        const Real* base = c.blob + offA[j];
#pragma unroll 1
        for (int i = 0; i < cnt; ++i) {
            __syncwarp();
#pragma unroll
            for (int w0 = 0; w0 < LD; w0 += 32) {
                const int w = w0 + lane;
                if (w < LD) W[w] = (w < n) ? base[i * n + w] : Real(0.0);
            }
            __syncwarp();
            rr_reflect<Real, NP, RPL>(t, W, Real(1.0));
        }
    }
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        if (row < m) {
            V2* dst = reinterpret_cast<V2*>(c.M + row * LD);
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                V2 v = t[s][k];
                if (2 * k < z0) v.x = Real(0.0);
                if (2 * k + 1 < z0) v.y = Real(0.0);
                dst[k] = v;
            }
        }
    }
}

/*
 * A dense task with a dense weight (`W * task` with a matrix, AutoStack.h:127; iHQP.cpp:152 forms A'WA): W = C C' (Cholesky,
 * in place over the staged W, lane r owns row r) turns the level into the identity-weight task (C'A, C'b), which then takes
 * the LQ path like any other: H = A'WA = (C'A)'(C'A), g = -A'Wb = -(C'A)'(C'b), null(C'A) = null(A). A: m x n row-major at
 * stride n, b: m, W: m x m row-major, all in the staged (writable) blob; m <= 32. Returns false if W is not positive definite.
 */
template <typename Real>
__device__ __forceinline__ bool rr_whiten(int lane, Real* __restrict__ A, Real* __restrict__ b, Real* __restrict__ W, int m, int n)
{
    bool spd = true;
#pragma unroll 1
    for (int col = 0; col < m; ++col) {
        Real s = Real(0.0);
        if (lane >= col && lane < m) {
            s = W[lane * m + col];
#pragma unroll 1
            for (int k = 0; k < col; ++k) s = fma(-W[lane * m + k], W[col * m + k], s);
        }
        const Real d = __shfl_sync(FULL, s, col);
        if (!(d > Real(0.0))) { spd = false; break; }          /* warp-uniform */
        const Real inv = fast_rsqrt(d);
        __syncwarp();                                            /* row `col` has been read by every lane */
        if (lane >= col && lane < m) W[lane * m + col] = s * inv;   /* C[r][col]; r == col: sqrt(d) */
        __syncwarp();
    }
    if (!spd) return false;
    Real bi = Real(0.0);                                         /* b' = C' b */
    if (lane < m)
#pragma unroll 1
        for (int r = lane; r < m; ++r) bi = fma(W[r * m + lane], b[r], bi);
    __syncwarp();
    if (lane < m) b[lane] = bi;
    /* A' = C' A in place: row i needs the original rows r >= i only, so ascending i is safe; lane per column */
#pragma unroll 1
    for (int j = lane; j < n; j += 32)
#pragma unroll 1
        for (int i = 0; i < m; ++i) {
            Real acc = Real(0.0);
#pragma unroll 1
            for (int r = i; r < m; ++r) acc = fma(W[r * m + i], A[r * n + j], acc);
            A[i * n + j] = acc;
        }
    __syncwarp();
    return true;
}

/*
 * Householder sweep: pushes the row space of the m rows in c.M (absolute columns, zero outside
 * the live basis) out of the basis. Afterwards the compact factor c.Lc (row r at Lc + r*LDL,
 * column b <-> basis column e0 + b, e0 = z0 rounded down to even) holds [L 0], c.piv[r] is the
 * pivot number of row r (negative: linearly dependent row), the basis rows ZB have taken the
 * reflectors, and the live columns [z0 + rank, n) span the null space. Returns the rank.
 * Replaces adding the optimality equalities of a finished level to every later QP
 * (iHQP.cpp:164-170, 314-332); dependent rows (|residual| <= tol_rank |row|) are skipped, the
 * job of qpOASES' linear-independence test (QProblem.cpp:2847).
 *
 * Phase 1, LQ of M: lane r owns row r in registers (the shared-memory rows are dead from then
 * on and take the reflectors: V[zc] = row zc of c.M). Step r: the owner publishes its row as
 * V[zc]; lanes over columns move the finished columns (< pivot column jz) of the published
 * copy into the compact factor and sum |a|^2 over the rest; u = a - rho e_jz replaces a in
 * V[zc]; every row takes H = I - beta u u' (rr_reflect): finished entries stay where they are
 * because u is zero there.
 * Phase 2: every basis row (lane per row, registers) takes u_0 .. u_{rank-1} back to back.
 */
template <typename Real, int NP, int RPL>
__device__ __forceinline__ int rr_sweep(RC<Real>& c, typename Tr<Real>::V2 (&t)[RPL][NP], int m)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP;
    const int lane = c.lane, z0 = c.z0, n = c.n, e0 = c.z0 & ~1, LDL = c.LDL;
    const Real tol2 = (Real)(c.p->tol_rank * c.p->tol_rank);
    Real rn2[RPL];
    RPROF_T(ps0);
    __syncwarp();           /* the rows of M were written by column lanes */
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        rr_load_if<Real, NP>(t[s], c.M + row * LD, row < m);
        Real a0 = 0, a1 = 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) { a0 = fma(t[s][k].x, t[s][k].x, a0); a1 = fma(t[s][k].y, t[s][k].y, a1); }
        rn2[s] = a0 + a1;
    }
#pragma unroll 1
    for (int i = lane; i < m * LDL; i += 32) c.Lc[i] = Real(0.0);
    int zc = 0;
#pragma unroll 1
    for (int r = 0; r < m; ++r) {
        const int jz = z0 + zc;
        Real* Vz = c.M + zc * LD;
        __syncwarp();
        rr_publish<Real, NP, RPL>(t, lane, r, Vz);
        __syncwarp();
        Real part = Real(0.0);
#pragma unroll
        for (int w0 = 0; w0 < LD; w0 += 32) {
            const int w = w0 + lane;
            if (w < LD) {
                const Real a = Vz[w];
                if (w < jz) { if (w >= e0) c.Lc[r * LDL + w - e0] = a; Vz[w] = Real(0.0); }
                else part = fma(a, a, part);
            }
        }
        const Real nn = warp_sum<Real>(part);
        const Real a_p = Vz[min(jz, LD - 1)];
        const Real rn2o = __shfl_sync(FULL, rr_pick<Real, RPL>(rn2, r >> 5), r & 31);
        if (!(nn > tol2 * rn2o) || nn <= Real(0.0)) { /* dependent (or zero) row, or basis exhausted */
            if (lane == 0) c.piv[r] = -1;
            continue;
        }
        const Real rs = fast_rsqrt(nn), sq = nn * rs;
        const Real rho = (a_p > Real(0.0)) ? -sq : sq;
        const Real beta = rs * fast_rcp(sq + fabs(a_p));             /* 2 / |u|^2 = 1 / (|a| (|a| + |a_p|)) */
        __syncwarp();       /* every lane has read Vz[jz] */
        if (lane == 0) { Vz[jz] = a_p - rho; c.piv[r] = zc; RR_BETA(c)[zc] = beta; c.Lc[r * LDL + jz - e0] = rho; }
        __syncwarp();
        rr_reflect<Real, NP, RPL>(t, Vz, beta);
        ++zc;
    }
    const int rank = zc;
    __syncwarp();
    RPROF_T(ps1);
    /* phase 2: basis rows (x rows and Cin rows) take all reflectors */
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        rr_load_if<Real, NP>(t[s], c.ZB + row * LD, row < c.NR0);
    }
#pragma unroll 1
    for (int k = 0; k < rank; ++k) rr_reflect<Real, NP, RPL>(t, c.M + k * LD, RR_BETA(c)[k]);
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        if (row < c.NR0) rr_store<Real, NP>(c.ZB + row * LD, t[s]);
    }
    __syncwarp();
#ifdef OSB_RR_PROFILE
    { RPROF_T(ps2); if (lane == 0 && blockIdx.x == OSB_RR_PROFILE) { rr_prof_sw[0] = ps1 - ps0; rr_prof_sw[1] = ps2 - ps1; } }
#endif
    return rank;
}

struct RRLevelOut { int status; int iters; };

/* offset of column j of the packed working-set factor R (q x q upper, by columns) */
__device__ __forceinline__ int rr_col(int j) { return (j * (j + 1)) >> 1; }

/*
 * Goldfarb-Idnani dual active-set iteration on the box lo <= y <= hi, y = y0 + N z, one row of
 * N per lane and slot in registers (t). The single-QP job of BackEnd::solve
 * (QPOasesBackEnd.cpp:248-307); method after Goldfarb & Idnani 1983 as used by the reference's
 * eiQuadProg back-end (external/eiQuadProg-ext/include/eiquadprog.hpp:287-470). Working-set
 * member k occupies column z0 + k of N; free columns are [z0 + q, n).
 *
 * The working-set factor R (N1' N1 = R' R) is kept as its INVERSE W = R^-1 (packed by columns
 * in RR_RP): r = R^-1 d1 is then one independent multiply-add per member and lane instead of a
 * back substitution with a shuffle per member; appending member q extends W by the column
 * [-r / rho ; 1 / rho]; removing member k0 deletes row k0 of W and rotates the columns
 * k0 .. q-1 by the plane rotations that carry row k0 of W onto the last axis (they are the
 * transposes of the rotations that would re-triangularise R without its column k0). All
 * rotations of a removal are known at once from the prefix sums of that row, and the same
 * rotations act on columns z0+k0 .. z0+q-1 of every row of N (a branch-free walk over the
 * register row with identity rotations elsewhere).
 */
template <typename Real, int NP, int RPL>
__device__ __forceinline__ RRLevelOut rr_active_set(RC<Real>& c, typename Tr<Real>::V2 (&t)[RPL][NP], Real (&yv)[RPL],
                                                    int (&wh)[RPL], int max_iter)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP, RS = 32 * RPL;
    constexpr bool QWIDE = LD > 32;     /* the working set can hold more than 32 members: lanes own r_k and r_{k+32} */
    const int lane = c.lane, z0 = c.z0, n = c.n, nz = n - z0;
    const Real tol = (Real)c.p->tol_feas;
    RRLevelOut out{OSB_OK, 0};
#ifdef OSB_RR_PROFILE
    long long pg[8] = {};
#endif
    bool parked = false, have = false, refreshed = false, refresh_pending = false;
    int q = 0, bidx = -1;
    Real up = Real(0.0);
    Real lop[RPL], hip[RPL];
    int eqf[RPL];
    {
        /* rows of N = (E Z) F^-T into registers; unconstrained minimiser y = y_prev - N p;
         * |row|^2 of N (invariant under the orthogonal column updates of the loop) */
        const V2* cs2 = reinterpret_cast<const V2*>(RR_CSCALE(c));
        const V2* p2 = reinterpret_cast<const V2*>(RR_PV(c));
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int row = lane + 32 * s;
            const V2* z2 = reinterpret_cast<const V2*>(c.ZB + ((row < c.NR) ? row : 0) * LD);   /* rowless lanes shadow row 0 */
            Real a0 = 0, a1 = 0, q0 = 0, q1 = 0;
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                if (RPL > 1 && k % 4 == 0 && k > 0) RR_FENCE();
                const V2 zv = z2[k], sc = cs2[k], pp = p2[k];
                V2 tv; tv.x = zv.x * sc.x; tv.y = zv.y * sc.y;
                t[s][k] = tv;
                a0 = fma(tv.x, pp.x, a0); a1 = fma(tv.y, pp.y, a1);
                q0 = fma(tv.x, tv.x, q0); q1 = fma(tv.y, tv.y, q1);
            }
            yv[s] = ((row < c.NR) ? RR_YPREV(c)[row] : Real(0.0)) - (a0 + a1);
            if (row < c.NR) RR_RQ(c)[row] = q0 + q1;
            wh[s] = -1;
            /* the box widened by the feasibility tolerance, once per level: pricing is then two
             * subtractions per row (infinite sides can never be violated) */
            const Real l = (row < c.NR) ? RR_LO(c)[row] : -Real(OSB_INFTY), h = (row < c.NR) ? RR_HI(c)[row] : Real(OSB_INFTY);
            lop[s] = (l > -Real(OSB_INFTY)) ? l - tol * fmax(Real(1.0), fabs(l)) : -Real(INFINITY);
            hip[s] = (h < Real(OSB_INFTY)) ? h + tol * fmax(Real(1.0), fabs(h)) : Real(INFINITY);
            eqf[s] = (l == h) ? 0x20000 : 0;
        }
    }
#pragma unroll 1
    for (;;) {
        RPROF_T(g0);
        if (!have) {
            /* ---- pricing: most violated row that is neither in the working set nor parked ---- */
            Real best = Real(0.0); bidx = -1;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int row = lane + 32 * s;
                const Real vl = lop[s] - yv[s], vh = yv[s] - hip[s];   /* > 0: violated beyond tolerance */
                const bool upper = vh > vl;
                const Real cand = upper ? vh : vl;
                if (wh[s] == -1 && cand > best) {       /* rowless lanes: infinite box, never a candidate */
                    best = cand;
                    bidx = row | (upper ? 0x10000 : 0) | eqf[s];
                }
            }
            bidx = warp_argmax_nonneg(best, bidx);
            if (!(best > Real(0.0))) break; /* primal feasible: optimal */
            up = Real(0.0);
            have = true;
        }
        if (++out.iters > max_iter) { out.status = OSB_MAX_ITER; return out; }
        const int cur = bidx & 0xffff;
        const Real csgn = (bidx & 0x10000) ? -Real(1.0) : Real(1.0); /* +1: lower side, -1: upper side */
        const int jq = z0 + q;
        RPROF_T(g1); RPROF_ADD(pg[0], g0, g1);
        /* d = N' n_p: the owner publishes its row; lanes over columns sign it (d), mask it to
         * the free columns (dm) and write the unit vector of the first free column (oh) */
        rr_publish<Real, NP, RPL>(t, lane, cur, RR_D(c));
        __syncwarp();
        Real zp = Real(0.0), mj[(LD + 31) / 32];
#pragma unroll
        for (int w0 = 0; w0 < LD; w0 += 32) {
            const int w = w0 + lane;
            mj[w0 / 32] = Real(0.0);
            if (w < LD) {
                const Real dj = csgn * RR_D(c)[w];
                const Real m_ = refresh_pending ? RR_VM(c)[w] : ((w >= jq) ? dj : Real(0.0));
                RR_D(c)[w] = dj; RR_DM(c)[w] = m_; RR_OH(c)[w] = (w == jq) ? Real(1.0) : Real(0.0);
                mj[w0 / 32] = m_;
                zp = fma(m_, m_, zp);
            }
        }
        __syncwarp();
        RPROF_T(g2); RPROF_ADD(pg[1], g1, g2);
        /*
         * Two independent strands in one basic block, for the scheduler to interleave:
         *  - scalars: |d2|^2 by shuffles, then rho, beta of the Householder v = d2 - rho e_q
         *    that an append would use (rsqrt, reciprocal: ~300 cycles of pure latency);
         *  - one walk over the register rows with two vectors: dy = N2 d2 and the column
         *    xq = N e_q of every row, which gives N2 v = dy - rho xq without waiting for rho.
         */
        const Real zz = warp_sum<Real>(zp);
        const Real dq = __shfl_sync(FULL, rr_pick<Real, (LD + 31) / 32>(mj, jq >> 5), jq & 31);
        const Real rs = fast_rsqrt(zz), sq = zz * rs;         /* 1/|d2|, |d2| */
        const Real rho = (dq > Real(0.0)) ? -sq : sq;
        const Real rho_inv = (dq > Real(0.0)) ? -rs : rs;
        const Real beta = rs * fast_rcp(sq + fabs(dq));   /* 1 / (rho (rho - dq)) */
        Real dyv[RPL], dvv[RPL];
        {
            const V2* d2 = reinterpret_cast<const V2*>(RR_DM(c));
            const V2* o2 = reinterpret_cast<const V2*>(RR_OH(c));
            Real ax[RPL], ay[RPL], bx[RPL], by[RPL];
#pragma unroll
            for (int s = 0; s < RPL; ++s) { ax[s] = ay[s] = Real(0.0); bx[s] = by[s] = Real(0.0); }
#pragma unroll
            for (int k = 0; k < NP; ++k) {
                if (k % (RPL > 1 ? (sizeof(Real) == 8 ? 2 : 3) : 5) == 0 && k > 0) RR_FENCE();
                const V2 a = d2[k], b = o2[k];
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    ax[s] = fma(t[s][k].x, a.x, ax[s]); ay[s] = fma(t[s][k].y, a.y, ay[s]);
                    bx[s] = fma(t[s][k].x, b.x, bx[s]); by[s] = fma(t[s][k].y, b.y, by[s]);
                }
            }
#pragma unroll
            for (int s = 0; s < RPL; ++s) { dyv[s] = ax[s] + ay[s]; dvv[s] = fma(-rho, bx[s] + by[s], dyv[s]); }
        }
        /* the Householder vector itself, for the walk of an append (made visible there) */
#pragma unroll
        for (int w0 = 0; w0 < LD; w0 += 32) {
            const int w = w0 + lane;
            if (w < LD) RR_VM(c)[w] = (w == jq) ? dq - rho : mj[w0 / 32];
        }
        if (__builtin_expect(refresh_pending, 0)) {      /* this pass was the refresh of a full vertex: dyv = N z */
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                const int row = lane + 32 * s;
                if (row < c.NR) yv[s] = RR_YPREV(c)[row] + dyv[s];
            }
            refresh_pending = false; refreshed = true; have = false;
            __syncwarp();
            continue;
        }
        RPROF_T(g3); RPROF_ADD(pg[2], g2, g3);
        /* r = W d1 (W = R^-1, upper triangular): lane k owns r_k (and r_{k+32}) */
        Real r0 = Real(0.0), r1 = Real(0.0);
        {
            Real r0b = Real(0.0);
            int j = 0;
#pragma unroll 1
            for (; j + 1 < q; j += 2) {
                const Real* colj = RR_RP(c) + rr_col(j);
                const Real* colk = colj + j + 1;
                const Real dj = RR_D(c)[z0 + j], dk = RR_D(c)[z0 + j + 1];
                const Real ea = (lane <= j) ? colj[lane] : Real(0.0);
                const Real eb = (lane <= j + 1) ? colk[lane] : Real(0.0);
                r0 = fma(ea, dj, r0); r0b = fma(eb, dk, r0b);
                if (QWIDE && j + 1 >= 32) {
                    if (lane + 32 <= j) r1 = fma(colj[lane + 32], dj, r1);
                    if (lane + 32 <= j + 1) r1 = fma(colk[lane + 32], dk, r1);
                }
            }
            if (j < q) {
                const Real* colj = RR_RP(c) + rr_col(j);
                const Real dj = RR_D(c)[z0 + j];
                if (lane <= j) r0 = fma(colj[lane], dj, r0);
                if (QWIDE && lane + 32 <= j) r1 = fma(colj[lane + 32], dj, r1);
            }
            r0 += r0b;
        }
        Real t1 = INFINITY; int lk = -1;
        if (q > 0) {
            /* ratio test over droppable (non-equality) working-set members */
#pragma unroll
            for (int h = 0; h < (QWIDE ? 2 : 1); ++h) {
                const int k = lane + 32 * h;
                const Real rk = h ? r1 : r0;
                if (k < q && rk > Real(0.0) && !(c.act[k] & 0x20000)) {
                    const Real ratio = fmax(RR_U(c)[k] * fast_rcp(rk), Real(0.0)); /* a multiplier is never negative */
                    if (ratio < t1) { t1 = ratio; lk = k; }
                }
            }
            lk = warp_argmin_nonneg(t1, lk);
        }
        RPROF_T(g4); RPROF_ADD(pg[3], g3, g4);
        const Real y_ip = __shfl_sync(FULL, rr_pick<Real, RPL>(yv, cur >> 5), cur & 31);
        const Real lim = (bidx & 0x10000) ? RR_HI(c)[cur] : RR_LO(c)[cur];
        const Real sv = csgn * (y_ip - lim); /* < 0 while violated */
        const bool indep = (zz > Tr<Real>::dep_tol * RR_RQ(c)[cur]) && (q < nz);
        const Real t2 = indep ? -sv * (rs * rs) : INFINITY;   /* -sv / |d2|^2 */
        const Real tt = fmin(t1, t2);
        if (__builtin_expect(!(tt < INFINITY), 0)) {
            /* neither a primal nor a dual step exists. At a FULL vertex (q == nz) the iterate is determined by the
             * working set alone: before believing the violation, y is recomputed from it, once per working set --
             * T z = bound - y_prev on the members (T = S R', so z = W'(S rhs): lane j dots column j of the inverse
             * factor), then y = y_prev + N z for every row. The running y += t dy carries the rounding of every step
             * times |row of N|^2 ~ 1/eps; after 60-100 changes on a heavily saturated robot stack that reached 7e-7
             * and turned a consistent vertex into a reported OSB_INFEASIBLE. */
            if (q == nz && !refreshed) {
                Real* v = RR_OH(c);
                Real* zr = RR_VM(c);
                __syncwarp();
#pragma unroll 1
                for (int h = 0; h < (QWIDE ? 2 : 1); ++h) {
                    const int k = lane + 32 * h;
                    if (k < q) {
                        const int a = c.act[k], row = a & 0xffff;
                        const Real limk = (a & 0x10000) ? RR_HI(c)[row] : RR_LO(c)[row];
                        v[k] = ((a & 0x10000) ? -Real(1.0) : Real(1.0)) * (limk - RR_YPREV(c)[row]);
                    }
                }
                __syncwarp();
#pragma unroll 1
                for (int w0 = 0; w0 < LD; w0 += 32) {
                    const int w = w0 + lane, j = w - z0;
                    if (w < LD) {
                        Real acc = Real(0.0);
                        if (j >= 0 && j < q) {
                            const Real* colj = RR_RP(c) + rr_col(j);
#pragma unroll 1
                            for (int i = 0; i <= j; ++i) acc = fma(colj[i], v[i], acc);
                        }
                        zr[w] = acc;
                    }
                }
                __syncwarp();
                /* the next pass of the loop is the refresh: its column phase feeds z to the walk instead of the masked
                 * candidate row, the walk's N z becomes y - y_prev (no second copy of the unrolled walk in the code) */
                refresh_pending = true;
                continue;
            }
            /* a row violated by rounding only at a degenerate vertex is parked until the iterate moves; anything
             * larger is a genuinely inconsistent constraint set (same rule as sot_kernel.cuh) */
            if (-sv <= Tr<Real>::degenerate_tol * fmax(Real(1.0), fabs(lim))) {
#pragma unroll
                for (int s = 0; s < RPL; ++s) if (lane + 32 * s == cur) wh[s] = -2;
                parked = true; have = false;
                __syncwarp();
                continue;
            }
#ifdef OSB_EMU_DEBUG
            if (lane == 0) printf("INFEASIBLE inst %d: iter %d q %d nz %d cur %d zz %.3e dd %.3e sv %.3e t1 %.3e indep %d\n", (int)blockIdx.x, out.iters, q, nz, cur, (double)zz, (double)RR_RQ(c)[cur], (double)sv, (double)t1, (int)indep);
#endif
            out.status = OSB_INFEASIBLE; return out;
        }
        if (indep) { /* primal + dual step */
#pragma unroll
            for (int s = 0; s < RPL; ++s) yv[s] = fma(tt, dyv[s], yv[s]);
        }
        if (lane < q) RR_U(c)[lane] = fma(-tt, r0, RR_U(c)[lane]);                                   /* lane k owns r_k, r_{k+32} */
        if (QWIDE && lane + 32 < q) RR_U(c)[lane + 32] = fma(-tt, r1, RR_U(c)[lane + 32]);
        up += tt;
        RPROF_T(g5); RPROF_ADD(pg[4], g3, g5);
        if (indep && t2 <= t1) {
            /* ---- full step: append `cur` as member q. Householder maps d2 onto rho e_q ---- */
            __syncwarp();       /* vm is complete */
            {
                const V2* v2 = reinterpret_cast<const V2*>(RR_VM(c));
                Real w[RPL];
#pragma unroll
                for (int s = 0; s < RPL; ++s) w[s] = beta * dvv[s];
#pragma unroll
                for (int k = 0; k < NP; ++k) {
                    if (RPL > 1 && k % (sizeof(Real) == 8 ? 2 : 4) == 0 && k > 0) RR_FENCE();
                    const V2 v = v2[k];
#pragma unroll
                    for (int s = 0; s < RPL; ++s) { t[s][k].x = fma(-w[s], v.x, t[s][k].x); t[s][k].y = fma(-w[s], v.y, t[s][k].y); }
                }
            }
            {
                Real* colq = RR_RP(c) + rr_col(q);
                if (lane < q) colq[lane] = -r0 * rho_inv;
                if (QWIDE && lane + 32 < q) colq[lane + 32] = -r1 * rho_inv;
                if (lane == 0) { colq[q] = rho_inv; c.act[q] = bidx; RR_U(c)[q] = up; }
            }
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                if (lane + 32 * s == cur) wh[s] = q;
                if (parked && wh[s] == -2) wh[s] = -1;   /* the iterate moved: parked rows are priced again */
            }
            parked = false; have = false; refreshed = false;
            ++q;
            __syncwarp();
            RPROF_T(g6); RPROF_ADD(pg[5], g5, g6);
            continue;
        }
        /* ---- partial (or pure dual) step: member k0 = lk leaves the working set ---- */
        {
            const int k0 = lk;
#pragma unroll
            for (int s = 0; s < RPL; ++s) {
                if (wh[s] == k0) wh[s] = -1;
                else if (wh[s] > k0) wh[s] -= 1;
            }
            __syncwarp();
            /* plane rotations from row k0 of W: rotation j (k0 <= j < q-1) acts on members j, j+1
             * with cs = w[j+1] / h[j+1], sn = h[j] / h[j+1], h[j]^2 = sum_{i=k0..j} w[i]^2.
             * (cs, sn) pairs are stored by absolute column, identity elsewhere. */
            V2* csn = reinterpret_cast<V2*>(RR_DM(c));      /* 2 LD entries: RR_DM and RR_VM */
#pragma unroll
            for (int w0 = 0; w0 < LD; w0 += 32) {
                const int w = w0 + lane;
                if (w < LD) { V2 id; id.x = Real(1.0); id.y = Real(0.0); csn[w] = id; }
            }
            __syncwarp();
            {
                Real carry = Real(0.0);                      /* h^2 up to the previous 32 members */
                const Real wk0 = RR_RP(c)[rr_col(k0) + k0];
#pragma unroll 1
                for (int base = 0; base < q; base += 32) {
                    const int j = base + lane;
                    const Real wj = (j >= k0 && j < q) ? RR_RP(c)[rr_col(j) + k0] : Real(0.0);
                    Real sc = wj * wj;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const Real up_ = __shfl_up_sync(FULL, sc, o);
                        if (lane >= o) sc += up_;
                    }
                    sc += carry;                              /* h[j]^2 */
                    Real prev = __shfl_up_sync(FULL, sc, 1);  /* h[j-1]^2 */
                    if (lane == 0) prev = carry;
                    if (j > k0 && j < q) {
                        /* the first rotation meets the signed w[k0], the later ones the positive h[j-1] */
                        const Real hinv = rsqrt(sc);
                        V2 rot; rot.x = wj * hinv; rot.y = ((j - 1 == k0) ? wk0 : sqrt(prev)) * hinv;
                        csn[z0 + j - 1] = rot;
                    }
                    carry = __shfl_sync(FULL, sc, 31);
                }
            }
            __syncwarp();
            /* W: delete row k0, rotate columns k0 .. q-1, drop the last column (lane per new row) */
            {
                Real ya[(RS + 31) / 32];                     /* running entry of my row(s) in column j */
#pragma unroll
                for (int s2 = 0; s2 < (RS + 31) / 32; ++s2) {
                    const int i = lane + 32 * s2;
                    ya[s2] = (i < k0 && i < q - 1) ? RR_RP(c)[rr_col(k0) + i] : Real(0.0);
                }
#pragma unroll 1
                for (int j = k0; j < q - 1; ++j) {
                    const V2 rot = csn[z0 + j];
                    Real yn[(RS + 31) / 32];
#pragma unroll
                    for (int s2 = 0; s2 < (RS + 31) / 32; ++s2) {
                        const int i = lane + 32 * s2, o = (i < k0) ? i : i + 1;   /* old row of new row i */
                        yn[s2] = (i < q - 1 && o <= j + 1) ? RR_RP(c)[rr_col(j + 1) + o] : Real(0.0);
                    }
                    __syncwarp();                             /* column j has been read by every lane */
#pragma unroll
                    for (int s2 = 0; s2 < (RS + 31) / 32; ++s2) {
                        const int i = lane + 32 * s2;
                        const Real a = ya[s2], b = yn[s2];
                        if (i <= j && i < q - 1) RR_RP(c)[rr_col(j) + i] = rot.x * a - rot.y * b;
                        ya[s2] = rot.y * a + rot.x * b;
                    }
                }
            }
            /* working-set bookkeeping: members above k0 move down by one */
            Real un = 0; int an = 0;
#pragma unroll 1
            for (int base = k0; base < q - 1; base += 32) {
                const int k = base + lane;
                if (k < q - 1) { un = RR_U(c)[k + 1]; an = c.act[k + 1]; }
                __syncwarp();
                if (k < q - 1) { RR_U(c)[k] = un; c.act[k] = an; }
                __syncwarp();
            }
            /* N: the same rotations on columns z0+k0 .. z0+q-1 of every row (identity elsewhere) */
#pragma unroll
            for (int cj = 0; cj < LD - 1; ++cj) {
                const V2 rot = csn[cj];
#pragma unroll
                for (int s = 0; s < RPL; ++s) {
                    Real& ea = (cj & 1) ? t[s][cj >> 1].y : t[s][cj >> 1].x;
                    Real& eb = ((cj + 1) & 1) ? t[s][(cj + 1) >> 1].y : t[s][(cj + 1) >> 1].x;
                    const Real a = ea, b = eb;
                    ea = rot.x * a - rot.y * b;
                    eb = rot.y * a + rot.x * b;
                }
            }
            --q;
            refreshed = false;
            __syncwarp();
        }
    }
#ifdef OSB_RR_PROFILE
    if (lane == 0 && blockIdx.x == OSB_RR_PROFILE) for (int i = 0; i < 6; ++i) rr_prof_gi[i] = pg[i];
#endif
    return out;
}

/* WH: the instance also carries rr_whiten (stacks with a dense weight on a dense task); the instances without it are the
 * ones every BASELINE configuration runs on (the extra code cost them 2 %: 0.184 -> 0.188 ms on C29 x 4096) */
template <typename Real, int NP, int RPL, bool WH = false>
__device__ __forceinline__ void rr_solve_impl(const sot_params& p)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP, RS = 32 * RPL;
    constexpr int NCP = 8, NC = 2 * NCP;   /* widest leading block of a level (register pairs / columns) */
    constexpr int NCS = NC + 2;            /* its row stride in shared memory: an odd number of 16-byte units (fp64), so that a lane walking
                                            * its own row of the reduced Hessian does not share a bank group with its seven neighbours */
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
    c.blob = ws + W.off_blob - W.stage_from; c.ZB = ws + W.off_ZB; c.M = ws + W.off_M; c.Lc = ws + W.off_L; c.LDL = W.LDL;
    c.vec = ws + W.off_vec; c.rowv = ws + W.off_rowv;
    c.act = (int32_t*)(ws + W.off_act); c.piv = (int32_t*)(ws + W.off_piv);
    const int stride_e = (sizeof(Real) == 8) ? L.stride : ((L.stride + 3) & ~3);
    Real* const out_x = static_cast<Real*>(p.x);
    Real* const out_xl = static_cast<Real*>(p.x_levels);
    const int n = D.n, cin = D.c_in;
    c.n = n; c.NR0 = n + cin; c.NR = c.NR0; c.z0 = 0;

    V2 t[RPL][NP];
    Real yv[RPL];
    int wh[RPL];

    /* ---- stage the level blocks of the instance blob: one TMA bulk copy; the global block
     * (bound sets, Cin, Ceq: each word is read exactly once, by consecutive lanes) is read
     * straight from global memory. Initialise the basis meanwhile. ---- */
    const Real* const gb = static_cast<const Real*>(p.blobs) + (size_t)inst * stride_e;
    const int g0 = W.stage_from;          /* first staged element (16-byte aligned) */
    blob_stage(ws + W.off_blob, gb + g0, (uint32_t)(stride_e - g0) * (uint32_t)sizeof(Real), ws + W.off_bar, lane);
#pragma unroll 1
    for (int i = lane; i < W.NRmax * NP; i += 32) {
        V2 zero; zero.x = Real(0.0); zero.y = Real(0.0);
        reinterpret_cast<V2*>(c.ZB)[i] = zero;
    }
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < W.NRmax; i += 32) {
        if (i < n) c.ZB[i * LD + i] = Real(1.0);
        RR_YPREV(c)[i] = Real(0.0);
    }
    const Real* blob = c.blob;

    /* ---- while the bulk copy is in flight: everything that needs the global block only. Its loads
     * go straight to global memory (~700 cycles each), so every loop below keeps several of them
     * in flight (rolled one-load loops cost 4 % of a 29-DoF stack and 15 % of the 38-DoF stack
     * with 26 global rows). ----
     * bounds: l = max l_s, u = min u_s, clipped to +-1e20; global rows lA <= Cin x <= uA */
    bool bad = false;       /* a NaN anywhere in the instance poisons everything downstream */
#pragma unroll
    for (int s = 0; s < RPL; ++s) {
        const int row = lane + 32 * s;
        Real l = -Real(OSB_INFTY), h = Real(OSB_INFTY);
        if (row < n) {
            Real lv[OSB_MAX_BOUND_SETS], hv[OSB_MAX_BOUND_SETS];
#pragma unroll
            for (int b = 0; b < OSB_MAX_BOUND_SETS; ++b) {
                const bool on = b < D.n_bound_sets;
                lv[b] = on ? gb[L.l[b] + row] : -Real(OSB_INFTY);
                hv[b] = on ? gb[L.u[b] + row] : Real(OSB_INFTY);
            }
#pragma unroll
            for (int b = 0; b < OSB_MAX_BOUND_SETS; ++b) {
                bad |= (lv[b] != lv[b]) | (hv[b] != hv[b]);
                l = fmax(l, lv[b]); h = fmin(h, hv[b]);
            }
        } else if (row < n + cin) {
            l = gb[L.lA + row - n]; h = gb[L.uA + row - n];
            bad |= (l != l) | (h != h);
        }
        if (row < W.NRmax) { RR_LO(c)[row] = clip_inf<Real>(l); RR_HI(c)[row] = clip_inf<Real>(h); }
        yv[s] = Real(0.0); wh[s] = -1;
    }
    {   /* Cin rows into the basis block, four rows (eight loads) per step */
        constexpr int JP = (2 * NP + 31) / 32;
#pragma unroll 1
        for (int r = 0; r < cin; r += 4) {
            Real v[4][JP];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                for (int jp = 0; jp < JP; ++jp) {
                    const int j = lane + 32 * jp;
                    v[rr][jp] = (r + rr < cin && j < n) ? gb[L.Cin + (r + rr) * n + j] : Real(0.0);
                }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                for (int jp = 0; jp < JP; ++jp) {
                    const int j = lane + 32 * jp;
                    bad |= (v[rr][jp] != v[rr][jp]);
                    if (r + rr < cin && j < n) c.ZB[(n + r + rr) * LD + j] = v[rr][jp];
                }
        }
    }
    if (D.c_eq > 0) {       /* the equality block is read again by the sweep: only the NaN scan here */
        const int cnt = D.c_eq * n + D.c_eq;
#pragma unroll 8
        for (int i = lane; i < cnt; i += 32) { const Real v = gb[L.Ceq + i]; bad |= (v != v); }
    }
    blob_wait(ws + W.off_bar);

    int status = OSB_OK, total_iters = 0;
    {
#pragma unroll 4
        for (int i = g0 + lane; i < L.stride; i += 32) { const Real v = blob[i]; bad |= (v != v); }
        if (__any_sync(FULL, bad)) status = OSB_NAN_INPUT;
    }
    __syncwarp();
    bool x_zero = true;     /* x_prev == 0 so far: residuals are the right-hand sides */
    bool identity = true;   /* no reflector has touched the basis yet: A Z = A         */
    /* reflectors of the swept levels, kept for rr_project_refl (not with global equalities: their
     * block is not staged, nothing dead to keep the reflectors in) */
    /* compiled only into the instances that have the registers for it: the fp64 two-rows-per-lane
     * instances sit at 255 registers and lost 2-3 % to the extra code (measured) */
    constexpr bool REFL = (RPL == 1) || (sizeof(Real) == 4);
    const bool keep_refl = REFL && D.c_eq == 0;
    unsigned long long ranks = 0ull;
    int nrefl = 0, last_lq = -1;          /* last active level that projects a dense task block */
#pragma unroll 1
    for (int k = 0; k < D.n_levels; ++k)
        if ((p.active_mask & (1u << k)) && rr_level_path_of(&D, k) == SOT_PATH_LQ) last_lq = k;

    const int last_level = 31 - __clz((int)(p.active_mask & ((D.n_levels >= 32) ? 0xffffffffu : ((1u << D.n_levels) - 1u))));

    /* stage -1 (only if there are global equalities Ceq x = deq) eliminates them before any
     * level; stages 0.. are the priority levels. One body: the sweep has a single call site. */
#pragma unroll 1
    for (int k = (D.c_eq > 0) ? -1 : 0; k < D.n_levels && status == OSB_OK; ++k) {
        const bool eqs = k < 0;
        if (!eqs && !(p.active_mask & (1u << k))) continue;
        const int kk = eqs ? 0 : k;
        const Real eps = (Real)p.eps_level[kk];
        const int z0 = c.z0, nz = n - z0;
        const int cl = eqs ? 0 : D.c_loc[kk];
        c.NR = c.NR0 + cl;
        const int NR = c.NR;
        const int max_iter = p.max_iter > 0 ? p.max_iter : 10 * (NR + n) + 20;
        if (nz == 0) { /* nothing left to decide: the solution is inherited */
            if (p.store_levels && !eqs)
#pragma unroll 1
                for (int i = lane; i < n; i += 32)
                    out_xl[((size_t)inst * D.n_levels + k) * n + i] = RR_YPREV(c)[i];
            continue;
        }
        /* ---- local rows: limits and value at x_prev ---- */
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int row = lane + 32 * s;
            if (row >= c.NR0) {
                const bool in = row < NR;
                if (in) {
                    RR_LO(c)[row] = clip_inf<Real>(blob[L.lCl[kk] + row - c.NR0]);
                    RR_HI(c)[row] = clip_inf<Real>(blob[L.uCl[kk] + row - c.NR0]);
                }
            }
        }
        if (cl > 0) {
            rr_residual<Real>(c, blob + L.Cl[kk], n, nullptr, cl, RR_YPREV(c), RR_YPREV(c) + c.NR0);
#pragma unroll 1
            for (int r = lane; r < cl; r += 32) RR_YPREV(c)[c.NR0 + r] = -RR_YPREV(c)[c.NR0 + r];
        }
        /*
         * Two ways to factor the reduced Hessian Hz = F F' of the level (DENSE-path stacks are
         * solved by sot_kernel.cuh):
         *  TRIVIAL  A = I, W = I:  Hz = (1+eps) I, nothing to factor.
         *  LQ       dense A, diagonal W: rotate the basis FIRST with the Householder sweep of
         *           M = A Z  (M Q = [Lm 0]); in the rotated basis Hz = blkdiag(Lm'W Lm + eps I,
         *           eps I): only a rank x rank Cholesky, and the same sweep is the basis
         *           shrink the next level needs.
         * p (RR_PV(c), absolute columns) becomes F^-1 g_z, the rows of ZB become N = (E Z) F^-T.
         */
        const bool trivial = !eqs && rr_level_path_of(&D, kk) == SOT_PATH_TRIVIAL;
#ifdef OSB_RR_PROFILE
        long long prof[8] = {}, prof_last = clock64();
#endif
        const int e0 = z0 & ~1;   /* the leading block of a level is columns [e0, e0 + ncl) */
        int ncl = 0, rank = 0;
        Real tail_scale, gdiag = Real(1.0);   /* gdiag: G[a][a] of the lane's row of the reduced Hessian */
        if (trivial) {
            tail_scale = Real(1.0) / sqrt(Real(1.0) + eps);
#pragma unroll 1
            for (int i = lane; i < n; i += 32)
                RR_VM(c)[i] = (Real(1.0) + eps) * RR_YPREV(c)[i] - blob[L.b[kk] + i] + (D.has_c[kk] ? blob[L.c[kk] + i] : Real(0.0));
        } else {
            const int m = eqs ? D.c_eq : D.m[kk];
            const Real* A = eqs ? gb + L.Ceq : blob + L.A[kk];
            if constexpr (WH) {
                if (!eqs && D.weight_kind[kk] == OSB_W_DENSE) {  /* dense weight: whiten the task in place (identity weight from here on) */
                    if (!rr_whiten<Real>(lane, c.blob + L.A[kk], c.blob + L.b[kk], c.blob + L.w[kk], m, n)) { status = OSB_NOT_SPD | (k << 8); break; }
                }
            }
            /* instruction counts per lane, measured: ~105 per reflector against ~290 per four task
             * rows and 32 basis columns of the product with the basis */
            bool by_refl = false;
            if constexpr (REFL) {
                by_refl = keep_refl && !identity && nrefl * 105 + 100 < 290 * ((m + 3) >> 2) * ((LD + 31) / 32);
                if (by_refl) rr_project_refl<Real, NP, RPL>(c, t, A, m, kk, ranks, L.A);
            }
            if (!by_refl) rr_project<Real, LD>(c, A, n, m, c.M, identity);
            if (!eqs) {
                if (x_zero) {
#pragma unroll 1
                    for (int r = lane; r < m; r += 32) RR_RES(c)[r] = blob[L.b[kk] + r];
                } else {
                    rr_residual<Real>(c, A, n, blob + L.b[kk], m, RR_YPREV(c), RR_RES(c));
                }
            }
            RPROF_MARK(0);   /* project + residual */
            rank = rr_sweep<Real, NP, RPL>(c, t, m);
            RPROF_MARK(1);   /* sweep */
            identity = false;
            if (REFL && keep_refl && kk < last_lq) {    /* A of this level is dead: its first rank * n words keep sqrt(beta) u */
#pragma unroll 1
                for (int i = 0; i < rank; ++i) {
                    const Real sb = sqrt(RR_BETA(c)[i]);
                    for (int w = lane; w < n; w += 32) c.blob[L.A[kk] + i * n + w] = sb * c.M[i * LD + w];
                }
                ranks |= (unsigned long long)rank << (8 * kk);
                nrefl += rank;
                __syncwarp();
            }
            if (eqs) {
                /* particular solution: forward substitution on the pivot rows, consistency check
                 * on the dependent ones (all lanes redundantly, c_eq is small) */
                Real* gz = RR_PV(c);
                int npiv = 0;
#pragma unroll 1
                for (int r = 0; r < m; ++r) {
                    Real s = gb[L.deq + r];
                    const int pc = c.piv[r];
                    const int upto = (pc >= 0) ? pc : npiv;
                    if (pc >= 0) ++npiv;
#pragma unroll 1
                    for (int j = 0; j < upto; ++j) s -= c.Lc[r * c.LDL + j] * gz[j];
                    if (pc >= 0) { if (lane == 0) gz[pc] = s / c.Lc[r * c.LDL + pc]; }
                    else if (fabs(s) > Tr<Real>::eq_tol * fmax(Real(1.0), fabs(gb[L.deq + r]))) status = OSB_INFEASIBLE;
                    __syncwarp();
                }
#pragma unroll 1
                for (int i = lane; i < c.NR0; i += 32) RR_YPREV(c)[i] = rr_dot_dyn<Real>(c.ZB + i * LD, gz, rank);
                c.z0 = rank;
                x_zero = false;
                __syncwarp();
                continue;
            }
            ncl = rank + (z0 & 1); tail_scale = Real(1.0) / sqrt(eps);
            const Real* wd = (D.weight_kind[kk] == OSB_W_DIAG) ? blob + L.w[kk] : nullptr;
#pragma unroll 1
            for (int r = lane; r < m; r += 32) RR_RES(c)[r] *= (wd ? wd[r] : Real(1.0));   /* W res */
#pragma unroll 1
            for (int i = lane; i < n; i += 32)                                          /* eps x_prev + c */
                RR_VM(c)[i] = eps * RR_YPREV(c)[i] + (D.has_c[kk] ? blob[L.c[kk] + i] : Real(0.0));
            /* ---- G = Lm' W Lm + eps I on the leading block (compact, row stride NC, over the dead
             * reflectors): lane a accumulates row a in registers and stores its strictly lower part
             * (zeros elsewhere: the Cholesky recurrence below walks whole rows); the diagonal stays
             * in a register ---- */
            {
                V2 g[NCP];
#pragma unroll
                for (int j = 0; j < NCP; ++j) { g[j].x = Real(0.0); g[j].y = Real(0.0); }
                Real dacc = Real(0.0);
#pragma unroll 2
                for (int r = 0; r < m; ++r) {
                    const Real la = (lane < ncl) ? c.Lc[r * c.LDL + lane] : Real(0.0);
                    const Real sc = la * (wd ? wd[r] : Real(1.0));
                    const V2* l2 = reinterpret_cast<const V2*>(c.Lc + r * c.LDL);
                    dacc = fma(sc, la, dacc);
#pragma unroll
                    for (int j = 0; j < NCP; ++j) {
                        const V2 lv = l2[j];
                        g[j].x = fma(sc, lv.x, g[j].x); g[j].y = fma(sc, lv.y, g[j].y);
                    }
                }
                gdiag = (lane < ncl) ? dacc + eps : Real(1.0);
                if (lane < ncl) {
                    V2* g2 = reinterpret_cast<V2*>(c.M + lane * NCS);
#pragma unroll
                    for (int j = 0; j < NCP; ++j) {
                        V2 v = g[j];
                        if (2 * j >= lane) v.x = Real(0.0);
                        if (2 * j + 1 >= lane) v.y = Real(0.0);
                        g2[j] = v;
                    }
                }
            }
        }
        __syncwarp();
        RPROF_MARK(2);       /* Gram */
        /* ---- g_z by basis column (lane per column): Z' v, minus Lm' W res on the leading block ---- */
        {
            const bool zpart = trivial || !x_zero || D.has_c[kk];
            const int m = D.m[kk];
#pragma unroll 1
            for (int j = lane; j < LD; j += 32) {
                Real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                if (j >= z0 && j < n) {
                    const Real* zc = c.ZB + j;
                    if (zpart) {
                        int i = 0;
#pragma unroll 1
                        for (; i + 3 < n; i += 4) {
                            s0 = fma(RR_VM(c)[i], zc[i * LD], s0); s1 = fma(RR_VM(c)[i + 1], zc[(i + 1) * LD], s1);
                            s2 = fma(RR_VM(c)[i + 2], zc[(i + 2) * LD], s2); s3 = fma(RR_VM(c)[i + 3], zc[(i + 3) * LD], s3);
                        }
#pragma unroll 1
                        for (; i < n; ++i) s0 = fma(RR_VM(c)[i], zc[i * LD], s0);
                    }
                    if (j < e0 + ncl) {
                        const Real* lc = c.Lc + (j - e0);
#pragma unroll 1
                        for (int r = 0; r + 1 < m; r += 2) {
                            s1 = fma(-RR_RES(c)[r], lc[r * c.LDL], s1); s2 = fma(-RR_RES(c)[r + 1], lc[(r + 1) * c.LDL], s2);
                        }
                        if (m & 1) s3 = fma(-RR_RES(c)[m - 1], lc[(m - 1) * c.LDL], s3);
                    }
                }
                RR_PV(c)[j] = (s0 + s1) + (s2 + s3);
            }
        }
        __syncwarp();
        RPROF_MARK(3);       /* g_z */
        /* ---- local rows in the (rotated) basis: extra rows of ZB ---- */
        if (cl > 0) { rr_project<Real, LD>(c, blob + L.Cl[kk], n, cl, c.ZB + c.NR0 * LD, false); __syncwarp(); }
        /*
         * ---- Cholesky G = F F' (left-looking, short shared-memory rows) with the gradient as
         * one more row: p = F^-1 g_z comes out of the same recurrence
         * X[c] = (X[c] - X[:c] . F[c][:c]) / F[c][c]; every lane carries the running diagonal
         * G[g][g] - |F[g][:c]|^2 of its row, so the pivot is one shuffle away. ----
         */
        bool spd = true;
        if (ncl > 0) {
            /* the gradient of the leading block becomes row ncl of the compact matrix (zero padded) */
            if (lane < NC) c.M[ncl * NCS + lane] = (lane < ncl) ? RR_PV(c)[e0 + lane] : Real(0.0);
            __syncwarp();
            Real dg = gdiag;                      /* G[g][g] - |F[g][:col]|^2 of my row */
            Real* const X = c.M + (lane <= ncl ? lane : 0) * NCS;   /* my row: G row, or the gradient */
            const V2* x2 = reinterpret_cast<const V2*>(X);
#pragma unroll 1
            for (int col = 0; col < ncl; ++col) {
                const V2* f2 = reinterpret_cast<const V2*>(c.M + col * NCS);
                const Real piv = __shfl_sync(FULL, dg, col);
                if (!(piv > Real(0.0))) { spd = false; break; }
                const Real inv = fast_rsqrt(piv);
                if (lane > col && lane <= ncl) {
                    Real a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                    for (int j = 0; j < NCP; j += 2) {
                        if (2 * j < col) {      /* warp-uniform: row `col` of F is zero from column col on */
                            const V2 xa = x2[j], xb = x2[j + 1];
                            const V2 fa = f2[j], fb = f2[j + 1];
                            a0 = fma(xa.x, fa.x, a0); a1 = fma(xa.y, fa.y, a1);
                            a2 = fma(xb.x, fb.x, a2); a3 = fma(xb.y, fb.y, a3);
                        }
                    }
                    const Real v = (X[col] - ((a0 + a1) + (a2 + a3))) * inv;
                    X[col] = v;
                    dg = fma(-v, v, dg);
                }
                if (lane == 0) RR_RINV(c)[col] = inv;
                __syncwarp();
            }
            if (lane < ncl) RR_PV(c)[e0 + lane] = c.M[ncl * NCS + lane];   /* p = F^-1 g_z on the leading block */
        }
        /*
         * ---- leading columns of N = (E Z) F^-T: every basis row runs the same recurrence on its
         * NC leading entries in registers (static indices, F rows broadcast), in place: those
         * columns leave the basis after this level ----
         */
        RPROF_MARK(4);       /* Cholesky */
        if (spd && ncl > 0) {
#pragma unroll 1
            for (int s = 0; s < RPL; ++s) {
                const int row = lane + 32 * s;
                if (row < NR) {
                    V2* z2 = reinterpret_cast<V2*>(c.ZB + row * LD + e0);
                    V2 x[NCP];
#pragma unroll
                    for (int j = 0; j < NCP; ++j) {
                        x[j].x = Real(0.0); x[j].y = Real(0.0);
                        if (e0 + 2 * j < LD) x[j] = z2[j];
                    }
#pragma unroll
                    for (int cc = 0; cc < NC; ++cc) {
                        if (cc < ncl) {
                            const V2* f2 = reinterpret_cast<const V2*>(c.M + cc * NCS);
                            Real a0 = (cc & 1) ? x[cc >> 1].y : x[cc >> 1].x, a1 = Real(0.0);
#pragma unroll
                            for (int j = 0; 2 * j < cc; ++j) {
                                const V2 fv = f2[j];
                                a0 = fma(-x[j].x, fv.x, a0);
                                if (2 * j + 1 < cc) a1 = fma(-x[j].y, fv.y, a1);
                            }
                            const Real v = (a0 + a1) * RR_RINV(c)[cc];
                            if (cc & 1) x[cc >> 1].y = v; else x[cc >> 1].x = v;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NCP; ++j) if (e0 + 2 * j < LD) z2[j] = x[j];
                }
            }
        }
        if (!spd) { status = OSB_NOT_SPD | (k << 8); break; }
        /* ---- column scales of the remaining columns, p on the tail ---- */
#pragma unroll 1
        for (int j = lane; j < LD; j += 32) {
            const bool live = (j >= z0) && (j < n);
            const bool lead = live && (j < e0 + ncl);
            RR_CSCALE(c)[j] = lead ? Real(1.0) : (live ? tail_scale : Real(0.0));
            if (!lead) RR_PV(c)[j] = live ? RR_PV(c)[j] * tail_scale : Real(0.0);
        }
        __syncwarp();
        /* ---- inequality rows and bounds ---- */
        RPROF_MARK(5);       /* leading columns of N, scales */
        RRLevelOut lo_out = rr_active_set<Real, NP, RPL>(c, t, yv, wh, max_iter);
        RPROF_MARK(6);       /* active-set loop */
#ifdef OSB_RR_PROFILE
        if (lane == 0 && blockIdx.x == OSB_RR_PROFILE)
            printf("L%d project %lld sweep %lld (ph1 %lld ph2 %lld) gram %lld gz %lld chol %lld lead+scales %lld GI %lld (%d iters)\n",
                   k, prof[0], prof[1], rr_prof_sw[0], rr_prof_sw[1], prof[2], prof[3], prof[4], prof[5], prof[6], lo_out.iters);
        if (lane == 0 && blockIdx.x == OSB_RR_PROFILE)
            printf("   GI: load/pricing %lld publish..vm %lld walk %lld matvec+ratio %lld (matvec..step %lld) append %lld\n",
                   rr_prof_gi[0], rr_prof_gi[1], rr_prof_gi[2], rr_prof_gi[3], rr_prof_gi[4], rr_prof_gi[5]);
#endif
        total_iters += lo_out.iters;
        if (lo_out.status != OSB_OK) { status = lo_out.status | (k << 8); break; }
        /* ---- level solution; active bounds are returned exactly on the bound ---- */
#pragma unroll
        for (int s = 0; s < RPL; ++s) {
            const int row = lane + 32 * s;
            if (row < c.NR0) {
                Real v = yv[s];
                if (row < n && wh[s] >= 0) v = (c.act[wh[s]] & 0x10000) ? RR_HI(c)[row] : RR_LO(c)[row];
                /* the box of the later levels contains this point: they keep it feasible in exact arithmetic (their z = 0),
                     * a row it misses by rounding (a parked row, <= degenerate_tol) would make their QP infeasible at a
                     * fully degenerate vertex (most bounds saturated): no-op whenever the point is inside */
                RR_YPREV(c)[row] = v;
                RR_LO(c)[row] = fmin(RR_LO(c)[row], v); RR_HI(c)[row] = fmax(RR_HI(c)[row], v);
                if (p.store_levels && row < n) out_xl[((size_t)inst * D.n_levels + k) * n + row] = v;
            }
        }
        x_zero = false;
        __syncwarp();
        /* ---- shrink the basis by the row space of this level's task ---- */
        if (!trivial) c.z0 += rank;
        else if (k != last_level) c.z0 = n;
    }
    /* ---- result ---- */
#pragma unroll 1
    for (int i = lane; i < n; i += 32) {
        const Real v = (status == OSB_NAN_INPUT) ? Real(0.0) : RR_YPREV(c)[i];
        out_x[(size_t)inst * n + i] = v;
        peer_store_x<Real>(p, inst, n, i, v);
    }
    if (lane == 0) {
        p.status[inst] = status;
        peer_store_status(p, inst, status);
        if (p.iters) p.iters[inst] = total_iters;
    }
}

#ifdef OSB_EMU
template <typename Real>
__device__ void rr_dispatch(const sot_params& p)
{
    if (p.rr.np == 15 && p.rr.rpl == 1) rr_solve_impl<Real, 15, 1, true>(p);
    else if (p.rr.np == 15 && p.rr.rpl == 2) rr_solve_impl<Real, 15, 2, true>(p);
    else rr_solve_impl<Real, 19, 2, true>(p);
}
#endif

} // namespace

#ifndef OSB_EMU
/* launch bounds: resident warps per SM the register budget must allow (shared memory is the
 * other limit; rr_plan_workspace sizes it) */
template <typename Real, int NP, int RPL, int MINB, bool WH = false>
__global__ void __launch_bounds__(32, MINB) rr_solve_kernel(const __grid_constant__ sot_params p)
{
    rr_solve_impl<Real, NP, RPL, WH>(p);
}
#endif

#endif

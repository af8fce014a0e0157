/*
 * sot_kernel_mw.cuh -- the hot path with register-resident rows and ONE CTA OF SEVERAL WARPS PER STACK.
 *
 * Same mathematics, reference semantics and status words as sot_kernel_rr.cuh (its header lists the
 * reference lines replaced: a-2, a-5, a-6, a-8, a-9 of SURVEY.md section 8a).
 * What is different is the mapping: stacks whose y = [x ; Cin x ; Cloc x] has more than 32 rows, or more
 * than 38 variables, used to be solved by one warp holding TWO rows per lane (255 registers, spills,
 * 6-8 resident warps per SM, a loop body twice as long as the instruction cache likes) or by the
 * all-in-shared-memory generic kernel at one warp per SM. Here thread t of a CTA of NW warps owns row t:
 * one row per lane like the 30x1 instance, NW times the resident warps for the same number of stacks
 * per SM, half the code per walk, and both warps of a stack run the same code at the same time.
 *
 * Division of labour (every choice keeps shared-memory writes single-writer between two barriers):
 *  - rows: thread tid owns row tid of whatever a phase transforms (basis rows during phase 2 of the
 *    sweep and the leading-column solve, rows of N in the active-set loop);
 *  - column-indexed work with stores (projection, g_z, scales): threads stride over columns;
 *  - small sequential pieces (phase 1 of a sweep: at most 32 task rows; Gram; Cholesky; the update of
 *    the packed inverse working-set factor) run on warp 0 with __syncwarp while the others wait at the
 *    next CTA barrier;
 *  - in the active-set loop each warp repeats the scalar decisions (ratio test, step length) from the same
 *    shared inputs and keeps PRIVATE copies of the vectors it derives from the published candidate row
 *    (signed row, masked row, Householder vector), so that an iteration needs three CTA barriers (pricing
 *    exchange, publication of the candidate row, before the common working-set state changes), four when
 *    a member leaves the working set.
 * Kernel instances: <Real, NP, NW> = register pairs per row, warps per CTA (mw_plan_workspace).
 */
#ifndef OSB_SOT_KERNEL_MW_CUH_
#define OSB_SOT_KERNEL_MW_CUH_

#include "sot_kernel_rr.cuh"

namespace {

#ifdef OSB_EMU
#define MW_SYNC() __syncthreads()
#else
#define MW_SYNC() __syncthreads()
#endif

/* row stride of the matrices a thread walks by rows (basis, task block): the 128-bit accesses of a quarter-warp fall on
 * eight different bank groups only if the stride is an ODD number of 16-byte units; 64 columns (NP = 32) put every row on the
 * same group (measured on C64c: 34 % of all shared-memory wavefronts were conflicts), two pad words cure it */
#define MW_LDR(NP) (((NP) & 1) ? 2 * (NP) : 2 * (NP) + 2)

template <typename Real>
struct MC {
    const sot_params* p;
    int lane, wid, tid;
    Real *blob, *ZB, *M, *Lc;
    Real *cvec, *pvec, *rowv, *scal;
    int32_t *act, *piv;
    int LDL;
    int n, NR0, NR;
    int z0;
};
/* common column-indexed vectors (LD entries each) */
#define MW_PV(c)     ((c).cvec)               /* gradient / p = F^-1 g_z                              */
#define MW_CSCALE(c) ((c).cvec + LD)          /* column scales of N                                   */
#define MW_U(c)      ((c).cvec + 2 * LD)      /* multipliers of the working set                       */
#define MW_RINV(c)   ((c).cvec + 3 * LD)      /* 1 / F[c][c]                                          */
#define MW_RAW(c)    ((c).cvec + 4 * LD)      /* published candidate row; task residual before that   */
#define MW_CSN(c)    ((c).cvec + 5 * LD)      /* 2 LD: rotations of a removal; before the loop:       */
#define MW_RES(c)    MW_RAW(c)
#define MW_BETA(c)   ((c).cvec + 5 * LD)      /*   reflector scales                                   */
#define MW_GV(c)     ((c).cvec + 6 * LD)      /*   eps x_prev + c, the operand of g_z                 */
/* private column-indexed vectors of a warp (4 LD per warp) */
#define MW_D(c)      ((c).pvec)               /* candidate row of N, signed                           */
#define MW_DM(c)     ((c).pvec + LD)          /* ... masked to the free columns                       */
#define MW_VM(c)     ((c).pvec + 2 * LD)      /* Householder vector of an append                      */
#define MW_OH(c)     ((c).pvec + 3 * LD)      /* unit vector of the first free column                 */
/* row-indexed vectors (RS = 32 NW entries each) */
#define MW_YPREV(c)  ((c).rowv)
#define MW_LO(c)     ((c).rowv + RS)
#define MW_HI(c)     ((c).rowv + 2 * RS)
#define MW_RQ(c)     ((c).rowv + 3 * RS)
#define MW_RP(c)     ((c).M)                  /* packed inverse working-set factor, over M            */
/* scalars: [0..2 NW) pricing slots (value, index) per warp, then y of the candidate row, rank */
#define MW_PRV(c, w) ((c).scal[2 * (w)])
#define MW_PRI(c, w) (reinterpret_cast<int32_t*>((c).scal + 2 * (w) + 1)[0])
#define MW_YIP(c)    ((c).scal[2 * NW])
#define MW_RANK(c)   (reinterpret_cast<int32_t*>((c).scal + 2 * NW + 1)[0])

/* dst rows <- A times the basis (or A itself while the basis is the identity): thread per basis column */
template <typename Real, int LD, int LDR, int NT>
__device__ __forceinline__ void mw_project(MC<Real>& c, const Real* __restrict__ A, int lda, int m, Real* dst, bool identity)
{
    const int n = c.n;
#pragma unroll 1
    for (int j = c.tid; j < LD; j += NT) {
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
                for (; i + 3 < n; i += 4) {
                    const Real* z = zcol + i * LDR;
                    const Real z0_ = z[0], z1_ = z[LDR], z2_ = z[2 * LDR], z3_ = z[3 * LDR];
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
                    const Real zv = zcol[i * LDR];
                    s0 = fma(a0[i], zv, s0); s1 = fma(a1[i], zv, s1);
                    s2 = fma(a2[i], zv, s2); s3 = fma(a3[i], zv, s3);
                }
            }
            dst[r * LDR + j] = s0; dst[r1 * LDR + j] = s1; dst[r2 * LDR + j] = s2; dst[r3 * LDR + j] = s3;
        }
    }
}
/* out[r] = rhs[r] - A[r,:] . x   (thread per row) */
template <typename Real, int NT>
__device__ __forceinline__ void mw_residual(MC<Real>& c, const Real* __restrict__ A, int lda, const Real* rhs, int m,
                                            const Real* x, Real* out)
{
#pragma unroll 1
    for (int r = c.tid; r < m; r += NT)
        out[r] = (rhs ? rhs[r] : Real(0.0)) - rr_dot_dyn<Real>(A + r * lda, x, c.n);
}

/*
 * Phase 1 of the Householder sweep (LQ of the m <= 32 rows in c.M), run by WARP 0 ALONE with warp-level
 * synchronisation: see rr_sweep in sot_kernel_rr.cuh for the step. Leaves the reflectors in the rows of
 * c.M, their scales in MW_BETA, the compact factor in c.Lc, pivots in c.piv; returns the rank.
 */
template <typename Real, int NP>
__device__ __forceinline__ int mw_sweep_lq(MC<Real>& c, typename Tr<Real>::V2 (&t)[1][NP], int m)
{
    constexpr int LD = 2 * NP, LDR = MW_LDR(NP);
    const int lane = c.lane, z0 = c.z0, e0 = c.z0 & ~1, LDL = c.LDL;
    const Real tol2 = (Real)(c.p->tol_rank * c.p->tol_rank);
    Real rn2;
    rr_load_if<Real, NP>(t[0], c.M + lane * LDR, lane < m);
    {
        Real a0 = 0, a1 = 0;
#pragma unroll
        for (int k = 0; k < NP; ++k) { a0 = fma(t[0][k].x, t[0][k].x, a0); a1 = fma(t[0][k].y, t[0][k].y, a1); }
        rn2 = a0 + a1;
    }
#pragma unroll 1
    for (int i = lane; i < m * LDL; i += 32) c.Lc[i] = Real(0.0);
    int zc = 0;
#pragma unroll 1
    for (int r = 0; r < m; ++r) {
        const int jz = z0 + zc;
        Real* Vz = c.M + zc * LDR;
        __syncwarp();
        if (lane == r) rr_store<Real, NP>(Vz, t[0]);
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
        const Real rn2o = __shfl_sync(FULL, rn2, r);
        if (!(nn > tol2 * rn2o) || nn <= Real(0.0)) {
            if (lane == 0) c.piv[r] = -1;
            continue;
        }
        const Real rs = fast_rsqrt(nn), sq = nn * rs;
        const Real rho = (a_p > Real(0.0)) ? -sq : sq;
        const Real beta = rs * fast_rcp(sq + fabs(a_p));
        __syncwarp();
        if (lane == 0) { Vz[jz] = a_p - rho; c.piv[r] = zc; MW_BETA(c)[zc] = beta; c.Lc[r * LDL + jz - e0] = rho; }
        __syncwarp();
        rr_reflect<Real, NP, 1>(t, Vz, beta, jz);
        ++zc;
    }
    __syncwarp();
    return zc;
}

/* Goldfarb-Idnani dual active set on the box of y, one row of N per thread (see rr_active_set) */
template <typename Real, int NP, int NW>
__device__ __forceinline__ RRLevelOut mw_active_set(MC<Real>& c, typename Tr<Real>::V2 (&t)[1][NP], Real& yv, int& wh, int max_iter)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP, LDR = MW_LDR(NP), RS = 32 * NW, NT = 32 * NW;
    constexpr bool QWIDE = LD > 32;
    const int lane = c.lane, wid = c.wid, tid = c.tid, z0 = c.z0, n = c.n, nz = n - z0;
    const Real tol = (Real)c.p->tol_feas;
    /* every vector a walk of this loop meets is zero on columns [0, z0), and so are the rows of N (their column scales are
     * zero there): when the first KS pairs lie there the walks enter behind them (one warp-uniform branch, exact zeros
     * skipped, bit-identical). fp64 only: measured -0.6 % (C38c) .. -6 % (C64c) there, +0.5 .. 2 % in fp32 */
    constexpr int KS = (sizeof(Real) == 8 && NP > 2 * RR_KS_AS) ? RR_KS_AS : 0;
    const bool full = (KS == 0) || z0 < 2 * KS;                  /* (see rr_active_set / rr_reflect) */
    RRLevelOut out{OSB_OK, 0};
    bool parked = false, have = false, refreshed = false, refresh_pending = false;
    int q = 0, bidx = -1;
    Real up = Real(0.0);
    Real lop, hip;
    int eqf;
    const bool has_row = tid < c.NR;
    {
        const V2* cs2 = reinterpret_cast<const V2*>(MW_CSCALE(c));
        const V2* p2 = reinterpret_cast<const V2*>(MW_PV(c));
        Real a0 = 0, a1 = 0, q0 = 0, q1 = 0;
        if (has_row) {
            const V2* z2 = reinterpret_cast<const V2*>(c.ZB + tid * LDR);
#define MW_AS_LOAD(k)                                                                                  \
            {                                                                                          \
                const V2 zv = z2[k], sc = cs2[k], pp = p2[k];                                          \
                V2 tv; tv.x = zv.x * sc.x; tv.y = zv.y * sc.y;                                         \
                t[0][k] = tv;                                                                          \
                a0 = fma(tv.x, pp.x, a0); a1 = fma(tv.y, pp.y, a1);                                    \
                q0 = fma(tv.x, tv.x, q0); q1 = fma(tv.y, tv.y, q1);                                    \
            }
            if (full) {
#pragma unroll
                for (int k = 0; k < KS; ++k) MW_AS_LOAD(k)
            } else {
#pragma unroll
                for (int k = 0; k < KS; ++k) { t[0][k].x = Real(0.0); t[0][k].y = Real(0.0); }
            }
#pragma unroll
            for (int k = KS; k < NP; ++k) MW_AS_LOAD(k)
#undef MW_AS_LOAD
        } else {
#pragma unroll
            for (int k = 0; k < NP; ++k) { t[0][k].x = Real(0.0); t[0][k].y = Real(0.0); }
        }
        yv = (has_row ? MW_YPREV(c)[tid] : Real(0.0)) - (a0 + a1);
        if (has_row) MW_RQ(c)[tid] = q0 + q1;
        wh = -1;
        const Real l = has_row ? MW_LO(c)[tid] : -Real(OSB_INFTY), h = has_row ? MW_HI(c)[tid] : Real(OSB_INFTY);
        lop = (l > -Real(OSB_INFTY)) ? l - tol * fmax(Real(1.0), fabs(l)) : -Real(INFINITY);
        hip = (h < Real(OSB_INFTY)) ? h + tol * fmax(Real(1.0), fabs(h)) : Real(INFINITY);
        eqf = (l == h) ? 0x20000 : 0;
    }
    MW_SYNC();      /* RQ complete; nothing of the previous phase is read any more */
#pragma unroll 1
    for (;;) {
        if (!have) {
            /* ---- pricing: most violated row that is neither in the working set nor parked ---- */
            Real best = Real(0.0); bidx = -1;
            {
                const Real vl = lop - yv, vh = yv - hip;
                const bool upper = vh > vl;
                const Real cand = upper ? vh : vl;
                if (wh == -1 && cand > best) { best = cand; bidx = tid | (upper ? 0x10000 : 0) | eqf; }
            }
            bidx = warp_argmax_nonneg(best, bidx);
            if (lane == 0) { MW_PRV(c, wid) = best; MW_PRI(c, wid) = bidx; }
            MW_SYNC();                                          /* barrier 1: pricing exchange */
            best = MW_PRV(c, 0); bidx = MW_PRI(c, 0);
#pragma unroll
            for (int w = 1; w < NW; ++w) {
                const Real bw = MW_PRV(c, w);
                if (bw > best) { best = bw; bidx = MW_PRI(c, w); }      /* ties: the lower row, as one warp would */
            }
            if (!(best > Real(0.0))) break; /* primal feasible: optimal */
            up = Real(0.0);
            have = true;
        }
        if (++out.iters > max_iter) { out.status = OSB_MAX_ITER; return out; }
        const int cur = bidx & 0xffff;
        const Real csgn = (bidx & 0x10000) ? -Real(1.0) : Real(1.0);
        const int jq = z0 + q;
        if (tid == cur) { rr_store<Real, NP>(MW_RAW(c), t[0]); MW_YIP(c) = yv; }
        MW_SYNC();                                              /* barrier 2: the candidate row is published */
        Real zp = Real(0.0), mj[(LD + 31) / 32];
#pragma unroll
        for (int w0 = 0; w0 < LD; w0 += 32) {
            const int w = w0 + lane;
            mj[w0 / 32] = Real(0.0);
            if (w < LD) {
                const Real dj = csgn * MW_RAW(c)[w];
                const Real m_ = refresh_pending ? MW_VM(c)[w] : ((w >= jq) ? dj : Real(0.0));
                MW_D(c)[w] = dj; MW_DM(c)[w] = m_; MW_OH(c)[w] = (w == jq) ? Real(1.0) : Real(0.0);
                mj[w0 / 32] = m_;
                zp = fma(m_, m_, zp);
            }
        }
        __syncwarp();
        const Real zz = warp_sum<Real>(zp);
        const Real dq = __shfl_sync(FULL, rr_pick<Real, (LD + 31) / 32>(mj, jq >> 5), jq & 31);
        const Real rs = fast_rsqrt(zz), sq = zz * rs;
        const Real rho = (dq > Real(0.0)) ? -sq : sq;
        const Real rho_inv = (dq > Real(0.0)) ? -rs : rs;
        const Real beta = rs * fast_rcp(sq + fabs(dq));
        Real dyv, dvv;
        {
            const V2* d2 = reinterpret_cast<const V2*>(MW_DM(c));
            const V2* o2 = reinterpret_cast<const V2*>(MW_OH(c));
            Real ax = 0, ay = 0, bx = 0, by = 0;
#define MW_AS_WALK(k)                                                                                  \
            {                                                                                          \
                if ((k) % 5 == 0 && (k) > 0) RR_FENCE();                                                \
                const V2 a = d2[k], b = o2[k];                                                         \
                ax = fma(t[0][k].x, a.x, ax); ay = fma(t[0][k].y, a.y, ay);                            \
                bx = fma(t[0][k].x, b.x, bx); by = fma(t[0][k].y, b.y, by);                            \
            }
            if (full) {
#pragma unroll
                for (int k = 0; k < KS; ++k) MW_AS_WALK(k)
            }
#pragma unroll
            for (int k = KS; k < NP; ++k) MW_AS_WALK(k)
#undef MW_AS_WALK
            dyv = ax + ay; dvv = fma(-rho, bx + by, dyv);
        }
#pragma unroll
        for (int w0 = 0; w0 < LD; w0 += 32) {
            const int w = w0 + lane;
            if (w < LD) MW_VM(c)[w] = (w == jq) ? dq - rho : mj[w0 / 32];
        }
        if (__builtin_expect(refresh_pending, 0)) {      /* this pass was the refresh of a full vertex: dyv = N z */
            if (has_row) yv = MW_YPREV(c)[tid] + dyv;
            refresh_pending = false; refreshed = true; have = false;
            MW_SYNC();          /* every warp has read RAW / YIP of this pass before the next publication */
            continue;
        }
        /* r = W d1 (W = R^-1, upper triangular): lane k owns r_k (and r_{k+32}); every warp for itself */
        Real r0 = Real(0.0), r1 = Real(0.0);
        {
            Real r0b = Real(0.0);
            int j = 0;
#pragma unroll 1
            for (; j + 1 < q; j += 2) {
                const Real* colj = MW_RP(c) + rr_col(j);
                const Real* colk = colj + j + 1;
                const Real dj = MW_D(c)[z0 + j], dk = MW_D(c)[z0 + j + 1];
                const Real ea = (lane <= j) ? colj[lane] : Real(0.0);
                const Real eb = (lane <= j + 1) ? colk[lane] : Real(0.0);
                r0 = fma(ea, dj, r0); r0b = fma(eb, dk, r0b);
                if (QWIDE && j + 1 >= 32) {
                    if (lane + 32 <= j) r1 = fma(colj[lane + 32], dj, r1);
                    if (lane + 32 <= j + 1) r1 = fma(colk[lane + 32], dk, r1);
                }
            }
            if (j < q) {
                const Real* colj = MW_RP(c) + rr_col(j);
                const Real dj = MW_D(c)[z0 + j];
                if (lane <= j) r0 = fma(colj[lane], dj, r0);
                if (QWIDE && lane + 32 <= j) r1 = fma(colj[lane + 32], dj, r1);
            }
            r0 += r0b;
        }
        Real t1 = INFINITY; int lk = -1;
        if (q > 0) {
#pragma unroll
            for (int h = 0; h < (QWIDE ? 2 : 1); ++h) {
                const int k = lane + 32 * h;
                const Real rk = h ? r1 : r0;
                if (k < q && rk > Real(0.0) && !(c.act[k] & 0x20000)) {
                    const Real ratio = fmax(MW_U(c)[k] * fast_rcp(rk), Real(0.0));
                    if (ratio < t1) { t1 = ratio; lk = k; }
                }
            }
            lk = warp_argmin_nonneg(t1, lk);
        }
        const Real y_ip = MW_YIP(c);
        const Real lim = (bidx & 0x10000) ? MW_HI(c)[cur] : MW_LO(c)[cur];
        const Real sv = csgn * (y_ip - lim);
        const bool indep = (zz > Tr<Real>::dep_tol * MW_RQ(c)[cur]) && (q < nz);
        const Real t2 = indep ? -sv * (rs * rs) : INFINITY;
        const Real tt = fmin(t1, t2);
        if (__builtin_expect(!(tt < INFINITY), 0)) {
            /* full vertex: recompute y from the working set before believing the violation (see rr_active_set); every
             * warp for its own rows, on its private vectors */
            if (q == nz && !refreshed) {
                Real* v = MW_OH(c);
                Real* zr = MW_VM(c);
                __syncwarp();
#pragma unroll 1
                for (int h = 0; h < (QWIDE ? 2 : 1); ++h) {
                    const int k = lane + 32 * h;
                    if (k < q) {
                        const int a = c.act[k], row = a & 0xffff;
                        const Real limk = (a & 0x10000) ? MW_HI(c)[row] : MW_LO(c)[row];
                        v[k] = ((a & 0x10000) ? -Real(1.0) : Real(1.0)) * (limk - MW_YPREV(c)[row]);
                    }
                }
                __syncwarp();
#pragma unroll 1
                for (int w0 = 0; w0 < LD; w0 += 32) {
                    const int w = w0 + lane, j = w - z0;
                    if (w < LD) {
                        Real acc = Real(0.0);
                        if (j >= 0 && j < q) {
                            const Real* colj = MW_RP(c) + rr_col(j);
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
                MW_SYNC();      /* every warp has read RAW / YIP of this pass before the owner publishes again */
                continue;
            }
            if (-sv <= Tr<Real>::degenerate_tol * fmax(Real(1.0), fabs(lim))) {
                if (tid == cur) wh = -2;
                parked = true; have = false;
                MW_SYNC();      /* every warp has read this candidate's slots before they are rewritten */
                continue;
            }
            out.status = OSB_INFEASIBLE; return out;
        }
        if (indep) yv = fma(tt, dyv, yv);
        up += tt;
        MW_SYNC();                                              /* barrier 3: every warp has read U, act, RP, RAW, YIP */
        if (wid == 0) {
            if (lane < q) MW_U(c)[lane] = fma(-tt, r0, MW_U(c)[lane]);
            if (QWIDE && lane + 32 < q) MW_U(c)[lane + 32] = fma(-tt, r1, MW_U(c)[lane + 32]);
        }
        if (indep && t2 <= t1) {
            /* ---- full step: append `cur` as member q ---- */
            __syncwarp();       /* the warp's vm is complete */
            {
                const V2* v2 = reinterpret_cast<const V2*>(MW_VM(c));
                const Real w = beta * dvv;
                if (full) {
#pragma unroll
                    for (int k = 0; k < KS; ++k) { const V2 v = v2[k]; t[0][k].x = fma(-w, v.x, t[0][k].x); t[0][k].y = fma(-w, v.y, t[0][k].y); }
                }
#pragma unroll
                for (int k = KS; k < NP; ++k) {
                    const V2 v = v2[k];
                    t[0][k].x = fma(-w, v.x, t[0][k].x); t[0][k].y = fma(-w, v.y, t[0][k].y);
                }
            }
            if (wid == 0) {
                Real* colq = MW_RP(c) + rr_col(q);
                if (lane < q) colq[lane] = -r0 * rho_inv;
                if (QWIDE && lane + 32 < q) colq[lane + 32] = -r1 * rho_inv;
                if (lane == 0) { colq[q] = rho_inv; c.act[q] = bidx; MW_U(c)[q] = up; }
            }
            if (tid == cur) wh = q;
            if (parked && wh == -2) wh = -1;
            parked = false; have = false; refreshed = false;
            ++q;
            continue;           /* the pricing barrier of the next iteration orders these writes before their readers */
        }
        /* ---- partial (or pure dual) step: member k0 = lk leaves the working set ---- */
        {
            const int k0 = lk;
            if (wh == k0) wh = -1;
            else if (wh > k0) wh -= 1;
            V2* csn = reinterpret_cast<V2*>(MW_CSN(c));      /* common: written by warp 0 */
            if (wid == 0) {
#pragma unroll
                for (int w0 = 0; w0 < LD; w0 += 32) {
                    const int w = w0 + lane;
                    if (w < LD) { V2 id; id.x = Real(1.0); id.y = Real(0.0); csn[w] = id; }
                }
                __syncwarp();
                Real carry = Real(0.0);
                const Real wk0 = MW_RP(c)[rr_col(k0) + k0];
#pragma unroll 1
                for (int base = 0; base < q; base += 32) {
                    const int j = base + lane;
                    const Real wj = (j >= k0 && j < q) ? MW_RP(c)[rr_col(j) + k0] : Real(0.0);
                    Real sc = wj * wj;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const Real up_ = __shfl_up_sync(FULL, sc, o);
                        if (lane >= o) sc += up_;
                    }
                    sc += carry;
                    Real prev = __shfl_up_sync(FULL, sc, 1);
                    if (lane == 0) prev = carry;
                    if (j > k0 && j < q) {
                        const Real hinv = rsqrt(sc);
                        V2 rot; rot.x = wj * hinv; rot.y = ((j - 1 == k0) ? wk0 : sqrt(prev)) * hinv;
                        csn[z0 + j - 1] = rot;
                    }
                    carry = __shfl_sync(FULL, sc, 31);
                }
                __syncwarp();
                /* W: delete row k0, rotate columns k0 .. q-1, drop the last column (lane per new row) */
                {
                    Real ya[(LD + 31) / 32];
#pragma unroll
                    for (int s2 = 0; s2 < (LD + 31) / 32; ++s2) {
                        const int i = lane + 32 * s2;
                        ya[s2] = (i < k0 && i < q - 1) ? MW_RP(c)[rr_col(k0) + i] : Real(0.0);
                    }
#pragma unroll 1
                    for (int j = k0; j < q - 1; ++j) {
                        const V2 rot = csn[z0 + j];
                        Real yn[(LD + 31) / 32];
#pragma unroll
                        for (int s2 = 0; s2 < (LD + 31) / 32; ++s2) {
                            const int i = lane + 32 * s2, o = (i < k0) ? i : i + 1;
                            yn[s2] = (i < q - 1 && o <= j + 1) ? MW_RP(c)[rr_col(j + 1) + o] : Real(0.0);
                        }
                        __syncwarp();
#pragma unroll
                        for (int s2 = 0; s2 < (LD + 31) / 32; ++s2) {
                            const int i = lane + 32 * s2;
                            const Real a = ya[s2], b = yn[s2];
                            if (i <= j && i < q - 1) MW_RP(c)[rr_col(j) + i] = rot.x * a - rot.y * b;
                            ya[s2] = rot.y * a + rot.x * b;
                        }
                    }
                }
                Real un = 0; int an = 0;
#pragma unroll 1
                for (int base = k0; base < q - 1; base += 32) {
                    const int k = base + lane;
                    if (k < q - 1) { un = MW_U(c)[k + 1]; an = c.act[k + 1]; }
                    __syncwarp();
                    if (k < q - 1) { MW_U(c)[k] = un; c.act[k] = an; }
                    __syncwarp();
                }
            }
            MW_SYNC();                                          /* barrier 4: the rotations are known to every warp */
            /* N: the same rotations on columns z0+k0 .. z0+q-1 of every row (identity elsewhere) */
#pragma unroll
            for (int cj = 0; cj < LD - 1; ++cj) {
                const V2 rot = csn[cj];
                Real& ea = (cj & 1) ? t[0][cj >> 1].y : t[0][cj >> 1].x;
                Real& eb = ((cj + 1) & 1) ? t[0][(cj + 1) >> 1].y : t[0][(cj + 1) >> 1].x;
                const Real a = ea, b = eb;
                ea = rot.x * a - rot.y * b;
                eb = rot.y * a + rot.x * b;
            }
            --q;
            refreshed = false;
            /* `have` stays set: the same candidate is published again (barrier 2 of the next iteration follows
             * this warp's reads of csn; warp 0 rewrites csn only after barrier 3 of that iteration) */
        }
    }
    return out;
}

template <typename Real, int NP, int NW, bool WH = false>      /* WH: see rr_solve_impl */
__device__ __forceinline__ void mw_solve_impl(const sot_params& p)
{
    typedef typename Tr<Real>::V2 V2;
    constexpr int LD = 2 * NP, LDR = MW_LDR(NP), RS = 32 * NW, NT = 32 * NW;
    constexpr int NCP = 8, NC = 2 * NCP, NCS = NC + 2;   /* NCS: row stride of the compact reduced Hessian (see sot_kernel_rr.cuh) */
    OSB_SMEM_RAW(ws_raw);
    Real* ws = reinterpret_cast<Real*>(ws_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int inst = blockIdx.x;
    if (inst >= p.batch) return;
    const osb_stack_desc& D = p.d;
    const osb_layout& L = p.L;
    const mw_workspace& W = p.mw;

    MC<Real> c;
    c.p = &p; c.lane = lane; c.wid = wid; c.tid = tid;
    c.blob = ws + W.off_blob - W.stage_from; c.ZB = ws + W.off_ZB; c.M = ws + W.off_M; c.Lc = ws + W.off_L; c.LDL = W.LDL;
    c.cvec = ws + W.off_cvec; c.pvec = ws + W.off_pvec + wid * 4 * LD; c.rowv = ws + W.off_rowv; c.scal = ws + W.off_scal;
    c.act = (int32_t*)(ws + W.off_act); c.piv = (int32_t*)(ws + W.off_piv);
    const int stride_e = (sizeof(Real) == 8) ? L.stride : ((L.stride + 3) & ~3);
    Real* const out_x = static_cast<Real*>(p.x);
    Real* const out_xl = static_cast<Real*>(p.x_levels);
    const int n = D.n, cin = D.c_in;
    c.n = n; c.NR0 = n + cin; c.NR = c.NR0; c.z0 = 0;

    V2 t[1][NP];
    Real yv = Real(0.0);
    int wh = -1;

    const Real* const gb = static_cast<const Real*>(p.blobs) + (size_t)inst * stride_e;
    const int g0 = W.stage_from;
    /* the global rows Cin go straight into their basis rows by the same engine, one bulk copy per row, when rows of n elements
     * keep 16-byte alignment on both sides (C38c, C64c in fp64): read through registers they cost four dependent round trips to
     * global memory at the start of every stack (5 % of all stall samples on C38c, profiles/r02_ncu_C38c_f64.txt) */
    const bool cin_tma = cin > 0 && (n * (int)sizeof(Real)) % 16 == 0 && (L.Cin * (int)sizeof(Real)) % 16 == 0;
    {
        const uint32_t blob_bytes = (uint32_t)(stride_e - g0) * (uint32_t)sizeof(Real), row_bytes = (uint32_t)n * (uint32_t)sizeof(Real);
        stage_begin(ws + W.off_bar, blob_bytes + (cin_tma ? (uint32_t)cin * row_bytes : 0u), tid);
        stage_copy(ws + W.off_blob, gb + g0, blob_bytes, ws + W.off_bar, tid);
        if (cin_tma)
#pragma unroll 1
            for (int r = 0; r < cin; ++r) stage_copy(c.ZB + (n + r) * LDR, gb + L.Cin + r * n, row_bytes, ws + W.off_bar, tid);
    }
#pragma unroll 1
    for (int i = tid; i < W.NRmax * (LDR / 2); i += NT) {
        const int row = i / (LDR / 2), col = 2 * (i - row * (LDR / 2));
        if (cin_tma && row >= n && row < n + cin && col < n) continue;      /* written by the bulk copy (n is even here) */
        V2 zero; zero.x = Real(0.0); zero.y = Real(0.0);
        reinterpret_cast<V2*>(c.ZB)[i] = zero;
    }
    MW_SYNC();
#pragma unroll 1
    for (int i = tid; i < W.NRmax; i += NT) {
        if (i < n) c.ZB[i * LDR + i] = Real(1.0);
        MW_YPREV(c)[i] = Real(0.0);
    }
    const Real* blob = c.blob;

    bool bad = false;
    {
        const int row = tid;
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
        if (row < W.NRmax) { MW_LO(c)[row] = clip_inf<Real>(l); MW_HI(c)[row] = clip_inf<Real>(h); }
    }
    if (!cin_tma) {   /* Cin rows into the basis block: four rows per step and warp */
        constexpr int JP = (2 * NP + 31) / 32;
#pragma unroll 1
        for (int r = 4 * wid; r < cin; r += 4 * NW) {
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
                    if (r + rr < cin && j < n) c.ZB[(n + r + rr) * LDR + j] = v[rr][jp];
                }
        }
    }
    if (D.c_eq > 0) {
        const int cnt = D.c_eq * n + D.c_eq;
#pragma unroll 4
        for (int i = tid; i < cnt; i += NT) { const Real v = gb[L.Ceq + i]; bad |= (v != v); }
    }
    blob_wait(ws + W.off_bar);

    int status = OSB_OK, total_iters = 0;
    {
#pragma unroll 4
        for (int i = g0 + tid; i < L.stride; i += NT) { const Real v = blob[i]; bad |= (v != v); }
        if (cin_tma)
#pragma unroll 1
            for (int r = 0; r < cin; ++r)
#pragma unroll 1
                for (int j = tid; j < n; j += NT) { const Real v = c.ZB[(n + r) * LDR + j]; bad |= (v != v); }
        if (__syncthreads_or(bad ? 1 : 0)) status = OSB_NAN_INPUT;
    }
    bool x_zero = true, identity = true;
    const int last_level = 31 - __clz((int)(p.active_mask & ((D.n_levels >= 32) ? 0xffffffffu : ((1u << D.n_levels) - 1u))));

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
        if (nz == 0) {
            if (p.store_levels && !eqs)
#pragma unroll 1
                for (int i = tid; i < n; i += NT)
                    out_xl[((size_t)inst * D.n_levels + k) * n + i] = MW_YPREV(c)[i];
            continue;
        }
        /* ---- local rows: limits and value at x_prev ---- */
        if (tid >= c.NR0 && tid < NR) {
            MW_LO(c)[tid] = clip_inf<Real>(blob[L.lCl[kk] + tid - c.NR0]);
            MW_HI(c)[tid] = clip_inf<Real>(blob[L.uCl[kk] + tid - c.NR0]);
        }
        if (cl > 0) {
#pragma unroll 1
            for (int r = tid; r < cl; r += NT)
                MW_YPREV(c)[c.NR0 + r] = rr_dot_dyn<Real>(blob + L.Cl[kk] + r * n, MW_YPREV(c), n);
        }
        const bool trivial = !eqs && rr_level_path_of(&D, kk) == SOT_PATH_TRIVIAL;
        const int e0 = z0 & ~1;
        int ncl = 0, rank = 0;
        Real tail_scale, gdiag = Real(1.0);
        if (trivial) {
            tail_scale = Real(1.0) / sqrt(Real(1.0) + eps);
#pragma unroll 1
            for (int i = tid; i < n; i += NT)
                MW_GV(c)[i] = (Real(1.0) + eps) * MW_YPREV(c)[i] - blob[L.b[kk] + i] + (D.has_c[kk] ? blob[L.c[kk] + i] : Real(0.0));
        } else {
            const int m = eqs ? D.c_eq : D.m[kk];
            const Real* A = eqs ? gb + L.Ceq : blob + L.A[kk];
            if constexpr (WH) {
                if (!eqs && D.weight_kind[kk] == OSB_W_DENSE) {  /* dense weight: warp 0 whitens the task in place (rr_whiten) */
                    bool ok = true;
                    if (wid == 0) ok = rr_whiten<Real>(lane, c.blob + L.A[kk], c.blob + L.b[kk], c.blob + L.w[kk], m, n);
                    if (__syncthreads_or(ok ? 0 : 1)) { status = OSB_NOT_SPD | (k << 8); break; }
                }
            }
            mw_project<Real, LD, LDR, NT>(c, A, n, m, c.M, identity);
            if (!eqs) {
                if (x_zero) {
#pragma unroll 1
                    for (int r = tid; r < m; r += NT) MW_RES(c)[r] = blob[L.b[kk] + r];
                } else {
                    mw_residual<Real, NT>(c, A, n, blob + L.b[kk], m, MW_YPREV(c), MW_RES(c));
                }
            }
            MW_SYNC();      /* the rows of M were written by column threads */
            /* phase 1: LQ of the (at most 32) task rows, warp 0 alone */
            if (wid == 0) {
                const int rk = mw_sweep_lq<Real, NP>(c, t, m);
                if (lane == 0) MW_RANK(c) = rk;
            }
            MW_SYNC();
            rank = MW_RANK(c);
            /* phase 2: every basis row (thread per row) takes all reflectors */
            if (tid < c.NR0) {
                rr_load<Real, NP>(t[0], c.ZB + tid * LDR);
#pragma unroll 1
                for (int kq = 0; kq < rank; ++kq) rr_reflect<Real, NP, 1>(t, c.M + kq * LDR, MW_BETA(c)[kq], c.z0 + kq);
                rr_store<Real, NP>(c.ZB + tid * LDR, t[0]);
            }
            identity = false;
            if (eqs) {
                MW_SYNC();
                /* particular solution: forward substitution on the pivot rows, consistency check on the dependent ones */
                Real* gz = MW_PV(c);
                int npiv = 0;
                if (wid == 0) {
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
                }
                status = __syncthreads_or(status == OSB_INFEASIBLE ? 1 : 0) ? OSB_INFEASIBLE : status;
#pragma unroll 1
                for (int i = tid; i < c.NR0; i += NT) MW_YPREV(c)[i] = rr_dot_dyn<Real>(c.ZB + i * LDR, gz, rank);
                c.z0 = rank;
                x_zero = false;
                MW_SYNC();
                continue;
            }
            ncl = rank + (z0 & 1); tail_scale = Real(1.0) / sqrt(eps);
            const Real* wd = (D.weight_kind[kk] == OSB_W_DIAG) ? blob + L.w[kk] : nullptr;
#pragma unroll 1
            for (int r = tid; r < m; r += NT) MW_RES(c)[r] *= (wd ? wd[r] : Real(1.0));
#pragma unroll 1
            for (int i = tid; i < n; i += NT)
                MW_GV(c)[i] = eps * MW_YPREV(c)[i] + (D.has_c[kk] ? blob[L.c[kk] + i] : Real(0.0));
            MW_SYNC();      /* reflectors in M are dead (phase 2 done), RES and GV complete */
            /* G = Lm' W Lm + eps I on the leading block, warp 0 (rows in registers, strictly lower part stored) */
            if (wid == 0) {
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
        if (trivial) MW_SYNC();     /* GV complete (the LQ branch synchronised before its Gram) */
        /* ---- g_z by basis column (thread per column): Z' v, minus Lm' W res on the leading block ---- */
        {
            const bool zpart = trivial || !x_zero || D.has_c[kk];
            const int m = D.m[kk];
#pragma unroll 1
            for (int j = tid; j < LD; j += NT) {
                Real s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                if (j >= z0 && j < n) {
                    const Real* zc = c.ZB + j;
                    if (zpart) {
                        int i = 0;
#pragma unroll 1
                        for (; i + 3 < n; i += 4) {
                            s0 = fma(MW_GV(c)[i], zc[i * LDR], s0); s1 = fma(MW_GV(c)[i + 1], zc[(i + 1) * LDR], s1);
                            s2 = fma(MW_GV(c)[i + 2], zc[(i + 2) * LDR], s2); s3 = fma(MW_GV(c)[i + 3], zc[(i + 3) * LDR], s3);
                        }
#pragma unroll 1
                        for (; i < n; ++i) s0 = fma(MW_GV(c)[i], zc[i * LDR], s0);
                    }
                    if (j < e0 + ncl) {
                        const Real* lc = c.Lc + (j - e0);
#pragma unroll 1
                        for (int r = 0; r + 1 < m; r += 2) {
                            s1 = fma(-MW_RES(c)[r], lc[r * c.LDL], s1); s2 = fma(-MW_RES(c)[r + 1], lc[(r + 1) * c.LDL], s2);
                        }
                        if (m & 1) s3 = fma(-MW_RES(c)[m - 1], lc[(m - 1) * c.LDL], s3);
                    }
                }
                MW_PV(c)[j] = (s0 + s1) + (s2 + s3);
            }
        }
        /* ---- local rows in the (rotated) basis: extra rows of ZB ---- */
        if (cl > 0) mw_project<Real, LD, LDR, NT>(c, blob + L.Cl[kk], n, cl, c.ZB + c.NR0 * LDR, false);
        MW_SYNC();
        /* ---- Cholesky G = F F' with the gradient as one more row, warp 0 (see sot_kernel_rr.cuh) ---- */
        bool spd = true;
        if (ncl > 0) {
            if (wid == 0) {
                if (lane < NC) c.M[ncl * NCS + lane] = (lane < ncl) ? MW_PV(c)[e0 + lane] : Real(0.0);
                __syncwarp();
                Real dg = gdiag;
                Real* const X = c.M + (lane <= ncl ? lane : 0) * NCS;
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
                            if (2 * j < col) {
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
                    if (lane == 0) MW_RINV(c)[col] = inv;
                    __syncwarp();
                }
                if (spd && lane < ncl) MW_PV(c)[e0 + lane] = c.M[ncl * NCS + lane];
            }
            spd = !__syncthreads_or(spd ? 0 : 1);
        }
        /* ---- leading columns of N = (E Z) F^-T: every basis row, in registers ---- */
        if (spd && ncl > 0 && tid < NR) {
            V2* z2 = reinterpret_cast<V2*>(c.ZB + tid * LDR + e0);
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
                    const Real v = (a0 + a1) * MW_RINV(c)[cc];
                    if (cc & 1) x[cc >> 1].y = v; else x[cc >> 1].x = v;
                }
            }
#pragma unroll
            for (int j = 0; j < NCP; ++j) if (e0 + 2 * j < LD) z2[j] = x[j];
        }
        if (!spd) { status = OSB_NOT_SPD | (k << 8); break; }
        /* ---- column scales of the remaining columns, p on the tail ---- */
#pragma unroll 1
        for (int j = tid; j < LD; j += NT) {
            const bool live = (j >= z0) && (j < n);
            const bool lead = live && (j < e0 + ncl);
            MW_CSCALE(c)[j] = lead ? Real(1.0) : (live ? tail_scale : Real(0.0));
            if (!lead) MW_PV(c)[j] = live ? MW_PV(c)[j] * tail_scale : Real(0.0);
        }
        MW_SYNC();
        /* ---- inequality rows and bounds ---- */
        RRLevelOut lo_out = mw_active_set<Real, NP, NW>(c, t, yv, wh, max_iter);
        total_iters += lo_out.iters;
        if (lo_out.status != OSB_OK) { status = lo_out.status | (k << 8); break; }
        /* ---- level solution; active bounds are returned exactly on the bound ---- */
        if (tid < c.NR0) {
            Real v = yv;
            if (tid < n && wh >= 0) v = (c.act[wh] & 0x10000) ? MW_HI(c)[tid] : MW_LO(c)[tid];
            MW_YPREV(c)[tid] = v;
            /* later levels keep this point feasible (see sot_kernel_rr.cuh): their box contains it */
            MW_LO(c)[tid] = fmin(MW_LO(c)[tid], v); MW_HI(c)[tid] = fmax(MW_HI(c)[tid], v);
            if (p.store_levels && tid < n) out_xl[((size_t)inst * D.n_levels + k) * n + tid] = v;
        }
        x_zero = false;
        MW_SYNC();
        if (!trivial) c.z0 += rank;
        else if (k != last_level) c.z0 = n;
    }
    /* ---- result ---- */
#pragma unroll 1
    for (int i = tid; i < n; i += NT) {
        const Real v = (status == OSB_NAN_INPUT) ? Real(0.0) : MW_YPREV(c)[i];
        out_x[(size_t)inst * n + i] = v;
        peer_store_x<Real>(p, inst, n, i, v);
    }
    if (tid == 0) {
        p.status[inst] = status;
        peer_store_status(p, inst, status);
        if (p.iters) p.iters[inst] = total_iters;
    }
}

#ifdef OSB_EMU
template <typename Real>
__device__ void mw_dispatch(const sot_params& p)
{
    if (p.mw.np == 19 && p.mw.nw == 2) mw_solve_impl<Real, 19, 2, true>(p);
    else if (p.mw.np == 32 && p.mw.nw == 2) mw_solve_impl<Real, 32, 2, true>(p);
    else mw_solve_impl<Real, 32, 3, true>(p);
}
#endif

} // namespace

#ifndef OSB_EMU
template <typename Real, int NP, int NW, int MINB, bool WH = false>
__global__ void __launch_bounds__(32 * NW, MINB) mw_solve_kernel(const __grid_constant__ sot_params p)
{
    mw_solve_impl<Real, NP, NW, WH>(p);
}
#endif

#endif

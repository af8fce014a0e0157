/*
 * sot_common.cuh -- pieces shared by the solver kernels (sot_kernel.cuh: generic shapes,
 * sot_kernel_rr.cuh: register-resident rows): precision traits, warp reductions, and the TMA
 * bulk copy (cp.async.bulk + mbarrier, SASS UBLKCP) that stages an instance blob.
 * With OSB_EMU defined the file is compiled by g++ for the warp emulator of tests/emu (test
 * infrastructure only); the staging then degenerates to a memcpy by lane 0.
 */
#ifndef OSB_SOT_COMMON_CUH_
#define OSB_SOT_COMMON_CUH_

#ifdef OSB_EMU
extern unsigned char* emu_smem;
#define OSB_SMEM_RAW(name) unsigned char* name = emu_smem
#else
#define OSB_SMEM_RAW(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace {

constexpr unsigned FULL = 0xffffffffu;
/* precision traits: vector pair type, reciprocal intrinsic, tolerances that scale with the
 * unit round-off (fp64 values are the ones DESIGN.md states; fp32 ones keep the same meaning) */
template <typename Real> struct Tr;
template <> struct Tr<double> {
    typedef double2 V2;
    static __device__ __forceinline__ double rcp(double x) { return __drcp_rn(x); }
    static constexpr double dep_tol = 1e-22;   /* |d2|^2 <= dep_tol |d|^2 : normal depends on working set */
    static constexpr double degenerate_tol = 1e-7; /* largest violation attributed to rounding at a vertex: the
                                                    * iterate carries 1/sqrt(eps) = 1.5e5 times the unit round-off
                                                    * per update; 1e-8 flagged 0.2 % of heavily saturated 35-DoF
                                                    * stacks (> 100 changes per level) as infeasible by 1.5-4e-8 */
    static constexpr double eq_tol = 1e-9;
};
template <> struct Tr<float> {
    typedef float2 V2;
    static __device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
    static constexpr float dep_tol = 1e-9f;
    static constexpr float degenerate_tol = 1e-3f;
    static constexpr float eq_tol = 1e-4f;
};

/* reciprocal to full precision from the hardware approximation and two Newton steps: about
 * 55 cycles of latency on B200 against 82 for __drcp_rn and 134 for a division (measured,
 * tools/lat_microbench.cu); used where a step length or a Householder scale is formed and a
 * last-bit rounding difference is irrelevant */
__device__ __forceinline__ double fast_rcp(double x)
{
#ifdef OSB_EMU
    return 1.0 / x;
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
}
/* 1/sqrt(x) to full precision from the hardware approximation (about 2^-22) and two Newton steps:
 * 9 instructions against ~15 for rsqrt(), which also treats denormals and infinities; callers
 * pass positive finite values or test the result's consumer (x <= 0 returns the raw approximation) */
__device__ __forceinline__ double fast_rsqrt(double x)
{
#ifdef OSB_EMU
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    if (x > 0.0) {
        double xy = x * y, e = fma(-xy, y, 1.0);
        y = fma(0.5 * y, e, y);
        xy = x * y; e = fma(-xy, y, 1.0);
        y = fma(0.5 * y, e, y);
    }
    return y;
#endif
}
__device__ __forceinline__ float fast_rsqrt(float x) { return rsqrt(x); }
__device__ __forceinline__ float fast_rcp(float x)
{
#ifdef OSB_EMU
    return 1.0f / x;
#else
    return __frcp_rn(x);
#endif
}

template <typename Real>
__device__ __forceinline__ Real warp_sum(Real v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
template <typename Real>
__device__ __forceinline__ Real clip_inf(Real v)
{
    return v < -Real(OSB_INFTY) ? -Real(OSB_INFTY) : (v > Real(OSB_INFTY) ? Real(OSB_INFTY) : v);
}

/* exact arg-max / arg-min of non-negative doubles over the warp with two 32-bit redux ops:
 * the bit pattern of a non-negative Real orders like an unsigned integer. Ties go to the
 * lowest lane. Returns the winner's `idx`; `v` becomes the extremum on every lane. */
__device__ __forceinline__ int warp_argmax_nonneg(double& v, int idx)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(FULL, hi);
    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    const unsigned win = __ballot_sync(FULL, hi == mh && lo == ml);
    v = __hiloint2double((int)mh, (int)ml);
    return __shfl_sync(FULL, idx, __ffs(win) - 1);
}
__device__ __forceinline__ int warp_argmin_nonneg(double& v, int idx)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(FULL, hi);
    const unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
    const unsigned win = __ballot_sync(FULL, hi == mh && lo == ml);
    v = __hiloint2double((int)mh, (int)ml);
    return __shfl_sync(FULL, idx, __ffs(win) - 1);
}
__device__ __forceinline__ int warp_argmax_nonneg(float& v, int idx)
{
    const unsigned key = __float_as_uint(v);
    const unsigned mk = __reduce_max_sync(FULL, key);
    const unsigned win = __ballot_sync(FULL, key == mk);
    v = __uint_as_float(mk);
    return __shfl_sync(FULL, idx, __ffs(win) - 1);
}
__device__ __forceinline__ int warp_argmin_nonneg(float& v, int idx)
{
    const unsigned key = __float_as_uint(v);
    const unsigned mk = __reduce_min_sync(FULL, key);
    const unsigned win = __ballot_sync(FULL, key == mk);
    v = __uint_as_float(mk);
    return __shfl_sync(FULL, idx, __ffs(win) - 1);
}
#ifndef OSB_EMU
/* ---- TMA bulk copy global -> shared, completion on an mbarrier ---------------------- */
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

#endif /* !OSB_EMU */

/* fused all-gather: the result of instance `inst` also goes, by plain peer stores over NVLink, into
 * the gather buffer of every connected GPU at global row peer_first + inst (no-op when the engine
 * was not launched through osb_solve_device_gather: n_peers == 0) */
template <typename Real>
__device__ __forceinline__ void peer_store_x(const sot_params& p, int inst, int n, int i, Real v)
{
#pragma unroll 1
    for (int q = 0; q < p.n_peers; ++q)
        static_cast<Real*>(p.peer_x[q])[(size_t)(p.peer_first + inst) * n + i] = v;
}
__device__ __forceinline__ void peer_store_status(const sot_params& p, int inst, int32_t status)
{
#pragma unroll 1
    for (int q = 0; q < p.n_peers; ++q) p.peer_status[q][p.peer_first + inst] = status;
}

/* one instance blob global -> shared: lane 0 issues a single bulk copy; completion is awaited
 * with blob_wait() so that independent start-up work overlaps the transfer */
__device__ __forceinline__ void blob_stage(void* dst, const void* src, uint32_t bytes, void* bar_word, int lane)
{
#ifdef OSB_EMU
    if (lane == 0) memcpy(dst, src, bytes);
    (void)bar_word;
#else
    const uint32_t bar = smem_u32(bar_word);
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(dst), src, bytes, bar);
    }
#endif
}
/* the same in pieces, for several bulk copies that complete on ONE mbarrier: stage_begin (thread `lane` == 0 initialises the
 * barrier and announces the total byte count), any number of stage_copy (thread 0 issues), then blob_wait by everybody.
 * Sizes and both addresses of every piece are multiples of 16 bytes. */
__device__ __forceinline__ void stage_begin(void* bar_word, uint32_t total_bytes, int lane)
{
#ifdef OSB_EMU
    (void)bar_word; (void)total_bytes; (void)lane;
#else
    if (lane == 0) { const uint32_t bar = smem_u32(bar_word); mbar_init(bar, 1); mbar_expect_tx(bar, total_bytes); }
#endif
}
__device__ __forceinline__ void stage_copy(void* dst, const void* src, uint32_t bytes, void* bar_word, int lane)
{
#ifdef OSB_EMU
    if (lane == 0) memcpy(dst, src, bytes);
    (void)bar_word;
#else
    if (lane == 0) bulk_g2s(smem_u32(dst), src, bytes, smem_u32(bar_word));
#endif
}
__device__ __forceinline__ void blob_wait(void* bar_word)
{
#ifndef OSB_EMU
    mbar_wait(smem_u32(bar_word), 0);
#else
    (void)bar_word;
#endif
    __syncwarp();
}

} // namespace
#endif

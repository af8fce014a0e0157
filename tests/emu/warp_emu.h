/*
 * warp_emu.h -- TEST INFRASTRUCTURE ONLY: runs one CUDA warp of the solver kernels on the CPU.
 *
 * The authoring container has no GPU, so the kernel sources (opensot_b200/csrc/*.cuh) are also
 * compiled by g++ with OSB_EMU defined and every lane of a warp run as a fibre (ucontext) of one
 * OS thread. Warp collectives (__syncwarp, shuffles, ballots, redux) are rendez-vous points: a
 * lane that reaches one yields; when all live lanes of ITS WARP have arrived the exchange happens and they
 * continue; CTA barriers (__syncthreads, __syncthreads_or) wait for every live thread of the CTA (the
 * CTA-per-stack kernel runs 2-3 warps). Between collectives lanes run one after another in a fixed order (ascending or
 * descending, chosen by the caller), so a cross-lane shared-memory dependency that is missing
 * its __syncwarp() shows up as a result that differs between the two orders.
 *
 * Nothing under opensot_b200/ links or loads this; the product path is CUDA only. The emulator is
 * used by tests/test_emu_kernel.py (marker: not gpu) to check the kernel logic against the oracle.
 */
#ifndef OSB_WARP_EMU_H_
#define OSB_WARP_EMU_H_

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <cmath>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(x) __attribute__((aligned(x)))

struct double2 { double x, y; } __attribute__((aligned(16)));
struct float2 { float x, y; } __attribute__((aligned(8)));
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }

struct emu_dim3 { int x, y, z; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace emu {
constexpr int W = 128;  /* threads of the largest emulated CTA (4 warps) */
struct State {
    ucontext_t sched;
    ucontext_t ctx[W];
    char* stacks[W];
    bool done[W];
    char waiting[W];    /* 0: running, 1: at a warp collective, 2: at a CTA barrier */
    uint64_t slot[W];   /* value deposited by a thread at the current collective */
    uint64_t snap[W];   /* values of the last completed collective               */
    int cur;
    int nthreads;       /* threads of the CTA being emulated (multiple of 32)    */
    int order;          /* 0: ascending threads, 1: descending                   */
    long collectives;
    uint64_t cta_or;    /* OR of the deposits of the last completed CTA barrier  */
};
extern State S;

/* deposit v, wait until every live lane of the WARP has arrived, then S.snap[] holds the warp's deposits */
static inline void rendezvous(uint64_t v)
{
    S.slot[S.cur] = v;
    S.waiting[S.cur] = 1;
    const int me = S.cur;
    swapcontext(&S.ctx[me], &S.sched);
    threadIdx.x = me;
}
/* the same for every live thread of the CTA (__syncthreads and its reducing variants) */
static inline void rendezvous_cta(uint64_t v)
{
    S.slot[S.cur] = v;
    S.waiting[S.cur] = 2;
    const int me = S.cur;
    swapcontext(&S.ctx[me], &S.sched);
    threadIdx.x = me;
}
template <typename T> static inline uint64_t to_bits(T v)
{
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> static inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
typedef void (*lane_fn)(void*);
void run_warp(lane_fn fn, void* arg, int order);
void run_cta(lane_fn fn, void* arg, int order, int nthreads);
}  // namespace emu

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::rendezvous(0); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{
    const int base = emu::S.cur & ~31;
    emu::rendezvous(emu::to_bits(v));
    return emu::from_bits<T>(emu::S.snap[base + (src & 31)]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m)
{
    const int me = emu::S.cur, base = me & ~31;
    emu::rendezvous(emu::to_bits(v));
    return emu::from_bits<T>(emu::S.snap[base + (((me & 31) ^ m) & 31)]);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d)
{
    const int me = emu::S.cur, l = me & 31;
    emu::rendezvous(emu::to_bits(v));
    return emu::from_bits<T>(emu::S.snap[(l + d) < 32 ? me + d : me]);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d)
{
    const int me = emu::S.cur, l = me & 31;
    emu::rendezvous(emu::to_bits(v));
    return emu::from_bits<T>(emu::S.snap[(l - d) >= 0 ? me - d : me]);
}
static inline unsigned __ballot_sync(unsigned, bool p)
{
    const int base = emu::S.cur & ~31;
    emu::rendezvous(p ? 1 : 0);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (!emu::S.done[base + i] && emu::S.snap[base + i]) r |= (1u << i);
    return r;
}
static inline void __syncthreads() { emu::rendezvous_cta(0); }
static inline int __syncthreads_or(int p) { emu::rendezvous_cta(p ? 1 : 0); return emu::S.cta_or ? 1 : 0; }
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, !p) == 0; }
static inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    const int base = emu::S.cur & ~31;
    emu::rendezvous(v);
    unsigned r = 0;
    for (int i = base; i < base + 32; ++i) if (!emu::S.done[i] && (unsigned)emu::S.snap[i] > r) r = (unsigned)emu::S.snap[i];
    return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
    const int base = emu::S.cur & ~31;
    emu::rendezvous(v);
    unsigned r = 0xffffffffu;
    for (int i = base; i < base + 32; ++i) if (!emu::S.done[i] && (unsigned)emu::S.snap[i] < r) r = (unsigned)emu::S.snap[i];
    return r;
}
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __double2hiint(double v) { return (int)(emu::to_bits(v) >> 32); }
static inline int __double2loint(double v) { return (int)(emu::to_bits(v) & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo)
{
    return emu::from_bits<double>(((uint64_t)(unsigned)hi << 32) | (uint64_t)(unsigned)lo);
}
static inline unsigned __float_as_uint(float v) { return (unsigned)emu::to_bits(v); }
static inline float __uint_as_float(unsigned v) { return emu::from_bits<float>((uint64_t)v); }
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float rsqrt(float x) { return 1.0f / sqrtf(x); }
static inline long long clock64() { return 0; }
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
using std::sqrt;
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

#endif

/*
 * emu_driver.cpp -- TEST INFRASTRUCTURE ONLY (see warp_emu.h): the solver kernels of
 * opensot_b200/csrc compiled for the CPU, one warp = 32 fibres, behind a ctypes-friendly
 * entry point that mirrors launch_solve() of osb_capi.cu. Never linked into the product.
 */
#define OSB_EMU 1
#include "warp_emu.h"

emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
unsigned char* emu_smem = nullptr;

namespace emu {
State S;
static lane_fn g_fn;
static void* g_arg;
static void trampoline(int lane)
{
    threadIdx.x = lane;
    g_fn(g_arg);
    S.done[lane] = true;
    S.cur = lane;
    swapcontext(&S.ctx[lane], &S.sched);
}
void run_warp(lane_fn fn, void* arg, int order) { run_cta(fn, arg, order, 32); }
void run_cta(lane_fn fn, void* arg, int order, int nthreads)
{
    constexpr size_t STACK = 512 * 1024;
    g_fn = fn; g_arg = arg;
    S.order = order;
    S.nthreads = nthreads;
    const int W = nthreads;
    for (int i = 0; i < W; ++i) {
        if (!S.stacks[i]) S.stacks[i] = (char*)malloc(STACK);
        S.done[i] = false; S.waiting[i] = 0;
        getcontext(&S.ctx[i]);
        S.ctx[i].uc_stack.ss_sp = S.stacks[i];
        S.ctx[i].uc_stack.ss_size = STACK;
        S.ctx[i].uc_link = &S.sched;
        makecontext(&S.ctx[i], (void (*)())trampoline, 1, i);
    }
    for (;;) {
        bool progressed = false;
        for (int k = 0; k < W; ++k) {
            const int i = order ? (W - 1 - k) : k;
            if (S.done[i] || S.waiting[i]) continue;
            S.cur = i; threadIdx.x = i;
            swapcontext(&S.sched, &S.ctx[i]);
            progressed = true;
        }
        bool all_done = true, released = false;
        for (int i = 0; i < W; ++i) if (!S.done[i]) all_done = false;
        if (all_done) break;
        /* warp collectives: complete for every warp whose live lanes have all arrived at one */
        for (int base = 0; base < W; base += 32) {
            bool any = false, all = true;
            for (int i = base; i < base + 32; ++i)
                if (!S.done[i]) { any = true; if (S.waiting[i] != 1) all = false; }
            if (any && all) {
                for (int i = base; i < base + 32; ++i) { S.snap[i] = S.slot[i]; S.waiting[i] = 0; }
                S.collectives++; released = true;
            }
        }
        if (!released) {    /* CTA barrier: every live thread of the CTA has arrived */
            bool all = true;
            for (int i = 0; i < W; ++i) if (!S.done[i] && S.waiting[i] != 2) all = false;
            if (all) {
                uint64_t o = 0;
                for (int i = 0; i < W; ++i) if (!S.done[i]) o |= S.slot[i];
                S.cta_or = o;
                for (int i = 0; i < W; ++i) { S.snap[i] = S.slot[i]; S.waiting[i] = 0; }
                S.collectives++; released = true;
            }
        }
        if (!released && !progressed) {
            fprintf(stderr, "warp emulator: deadlock (a thread exited or diverged around a collective)\n");
            abort();
        }
    }
}
}  // namespace emu

#include "sot_params.h"
#include "sot_kernel.cuh"
#include "sot_kernel_rr.cuh"
#include "sot_kernel_mw.cuh"

struct emu_call {
    const sot_params* p;
    int variant;
};
static void lane_entry(void* a)
{
    const emu_call* c = (const emu_call*)a;
    if (c->variant == 0) sot_solve_impl<double>(*c->p);
    else if (c->variant == 1) sot_solve_impl<float>(*c->p);
    else if (c->variant == 2) rr_dispatch<double>(*c->p);
    else if (c->variant == 3) rr_dispatch<float>(*c->p);
    else if (c->variant == 4) mw_dispatch<double>(*c->p);
    else mw_dispatch<float>(*c->p);
}

/* variant: 0 generic fp64, 1 generic fp32, 2 register-row fp64, 3 register-row fp32, 4 CTA-per-stack fp64, 5 fp32.
 * order: 0 ascending / 1 descending lane schedule. Returns 0, or -1 if the shape is not
 * eligible for the requested kernel. info[0] = shared-memory bytes per warp. */
extern "C" int emu_solve_ex(const osb_stack_desc* d, const void* blobs, int32_t batch, void* x,
                            int32_t* status, void* x_levels, int32_t* iters, const osb_options* opt,
                            uint32_t active_mask, int variant, int order, int64_t* info, const double* eps_level);
extern "C" int emu_solve(const osb_stack_desc* d, const void* blobs, int32_t batch, void* x,
                         int32_t* status, void* x_levels, int32_t* iters, const osb_options* opt,
                         uint32_t active_mask, int variant, int order, int64_t* info)
{
    return emu_solve_ex(d, blobs, batch, x, status, x_levels, iters, opt, active_mask, variant, order, info, nullptr);
}
extern "C" int emu_solve_ex(const osb_stack_desc* d, const void* blobs, int32_t batch, void* x,
                            int32_t* status, void* x_levels, int32_t* iters, const osb_options* opt,
                            uint32_t active_mask, int variant, int order, int64_t* info, const double* eps_level)
{
    osb_layout L;
    if (osb_blob_layout(d, &L) != OSB_SUCCESS) return -2;
    sot_params p;
    memset(&p, 0, sizeof(p));
    const int es = (variant & 1) ? 4 : 8;
    p.d = *d; p.L = L;
    if (es == 4 && p.d.eps < OSB_F32_MIN_EPS) p.d.eps = OSB_F32_MIN_EPS;
    for (int k = 0; k < OSB_MAX_LEVELS; ++k) p.eps_level[k] = eps_level ? eps_level[k] : p.d.eps;
    int64_t smem;
    int nthreads = 32;
    if (variant < 2) smem = sot_plan_workspace(d, &L, &p.w, es);
    else if (variant >= 4) {
        smem = mw_plan_workspace(d, &L, &p.mw, es);
        if (smem < 0) return -1;
        nthreads = 32 * p.mw.nw;
    } else {
        smem = rr_plan_workspace(d, &L, &p.rr, es);
        if (smem < 0) return -1;
    }
    if (info) info[0] = smem;
    p.max_iter = opt ? opt->max_iter : 0;
    p.store_levels = (opt && opt->store_levels && x_levels) ? 1 : 0;
    p.tol_feas = (opt && opt->tol_feas > 0) ? opt->tol_feas : (es == 8 ? 1e-12 : 1e-6);
    p.tol_rank = (opt && opt->tol_rank > 0) ? opt->tol_rank : (es == 8 ? 1e-11 : 1e-5);
    p.active_mask = active_mask ? active_mask : ((1u << d->n_levels) - 1u);
    p.batch = batch; p.blobs = blobs; p.x = x; p.status = status; p.x_levels = x_levels; p.iters = iters;
    void* raw = nullptr;
    if (posix_memalign(&raw, 128, (size_t)smem + 128)) return -3;
    emu_smem = (unsigned char*)raw;
    blockDim.x = nthreads; gridDim.x = batch;
    emu_call c{&p, variant};
    for (int b = 0; b < batch; ++b) {
        blockIdx.x = b;
        memset(emu_smem, getenv("EMU_POISON") ? 0xA5 : 0, (size_t)smem); /* poison: uninitialised reads show up */
        emu::run_cta(lane_entry, &c, order, nthreads);
    }
    free(raw);
    emu_smem = nullptr;
    return 0;
}

/*
 * osb_capi.cu -- host side of the C ABI declared in include/opensot_b200.h.
 * Thin by design: validates shapes, plans the per-warp shared-memory workspace, owns the
 * device buffers of one engine and launches the kernels of sot_kernel.cu. No CPU solve
 * path exists here: without a CUDA device every compute entry point fails loudly.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "opensot_b200.h"
#include "sot_params.h"

#include "sot_kernel.cuh"    /* generic kernel: any shape the ABI admits                            */
#include "sot_kernel_rr.cuh" /* register-row kernel: n <= 38, TRIVIAL/LQ levels (one translation  */
                             /* unit, no relocatable device code)                                  */
#include "sot_kernel_mw.cuh" /* CTA-per-stack kernel: 33..96 rows of y or 39..64 variables         */
#include "sot_kinematics.cuh" /* Cartesian / CoM producers from joint states                         */
#include <stdlib.h>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

/* instances of the register-row kernel: <Real, pairs per row, rows per lane, resident warps per SM
 * the register allocation must allow> */
#ifndef OSB_RR_MINB_F32
#define OSB_RR_MINB_F32 20   /* resident warps per SM the 30x1 fp32 instance is compiled for */
#endif
#ifndef OSB_MW_MINB_F64
#define OSB_MW_MINB_F64 6    /* resident CTAs per SM the 38x2-warp fp64 instance is compiled for */
#endif
#ifndef OSB_RR_MINB_F64
#define OSB_RR_MINB_F64 12   /* resident warps per SM the 30x1 fp64 instance is compiled for */
#endif
typedef void (*rr_kernel_fn)(const sot_params);
static rr_kernel_fn rr_pick_kernel(int es, int np, int rpl, bool wh)
{
    if (es == 8 && wh) {    /* stacks with a dense weight on a dense task: the instances that carry rr_whiten (fp64 only: the fp32
                             * engine refuses such stacks, OSB_E_PRECISION) */
        if (np == 15 && rpl == 1) return rr_solve_kernel<double, 15, 1, OSB_RR_MINB_F64, true>;
        if (np == 15 && rpl == 2) return rr_solve_kernel<double, 15, 2, 6, true>;
        if (np == 19 && rpl == 2) return rr_solve_kernel<double, 19, 2, 6, true>;
    }
    if (es == 8) {
        if (np == 15 && rpl == 1) return rr_solve_kernel<double, 15, 1, OSB_RR_MINB_F64>;
        if (np == 15 && rpl == 2) return rr_solve_kernel<double, 15, 2, 6>;
        if (np == 19 && rpl == 2) return rr_solve_kernel<double, 19, 2, 6>;
    } else {
        if (np == 15 && rpl == 1) return rr_solve_kernel<float, 15, 1, OSB_RR_MINB_F32>;
        if (np == 15 && rpl == 2) return rr_solve_kernel<float, 15, 2, 12>;
        if (np == 19 && rpl == 2) return rr_solve_kernel<float, 19, 2, 12>;
    }
    return nullptr;
}

/* instances of the CTA-per-stack kernel: <Real, pairs per row, warps per CTA, resident CTAs per SM the
 * register allocation must allow> */
static rr_kernel_fn mw_pick_kernel(int es, int np, int nw, bool wh)
{
    if (es == 8 && wh) {
        if (np == 19 && nw == 2) return mw_solve_kernel<double, 19, 2, OSB_MW_MINB_F64, true>;
        if (np == 32 && nw == 2) return mw_solve_kernel<double, 32, 2, 4, true>;
        if (np == 32 && nw == 3) return mw_solve_kernel<double, 32, 3, 2, true>;
    }
    if (es == 8) {
        if (np == 19 && nw == 2) return mw_solve_kernel<double, 19, 2, OSB_MW_MINB_F64>;
        if (np == 32 && nw == 2) return mw_solve_kernel<double, 32, 2, 4>;
        if (np == 32 && nw == 3) return mw_solve_kernel<double, 32, 3, 2>;
    } else {
        if (np == 19 && nw == 2) return mw_solve_kernel<float, 19, 2, 10>;
        if (np == 32 && nw == 2) return mw_solve_kernel<float, 32, 2, 6>;
        if (np == 32 && nw == 3) return mw_solve_kernel<float, 32, 3, 4>;
    }
    return nullptr;
}

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, const char* detail)
{
    snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
    return code;
}
#define CUDA_OK(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) return fail(OSB_E_CUDA, #call ": %s", cudaGetErrorString(e__)); \
    } while (0)

/* inside engine_create, after the engine exists: free it and its device buffers on failure */
#define CUDA_OK_E(call)                                                            \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            osb_engine_destroy(e);                                                 \
            return fail(OSB_E_CUDA, #call ": %s", cudaGetErrorString(e__));        \
        }                                                                          \
    } while (0)

#define OSB_PIPE_CHUNKS 8      /* solve_host splits a batch into at most this many chunks  */
#define OSB_PIPE_LAST_CHUNK 256 /* ... the last of which has about this many instances       */

/* one set of device buffers and events of a host-buffer call. Slot 0 is the engine's own set (the blocking entry points
 * use it); slot 1 exists once osb_solve_host_submit / osb_solve_states_submit have two calls in flight */
struct osb_pipe_slot {
    char* d_blobs; char* d_x; int32_t* d_status; char* d_xlev; int32_t* d_iters; double* s_q;
    char* d_wire;                   /* wire-format input of this slot (osb_wire_configure) */
    cudaEvent_t ev_h2d[OSB_PIPE_CHUNKS], ev_k[OSB_PIPE_CHUNKS], ev_end;
    int32_t ready;
    int32_t tag;                    /* what the constant part of d_blobs holds: OSB_TAG_* */
};
enum { OSB_TAG_NONE = 0, OSB_TAG_STATES = 1, OSB_TAG_WIRE = 2 };
#define OSB_WIRE_MAX_SEGMENTS 16

struct osb_engine {
    osb_stack_desc d;
    osb_layout L;
    sot_workspace w;
    rr_workspace rr;
    rr_kernel_fn rr_kernel;         /* non-null: the register-row kernel solves this shape */
    mw_workspace mw;
    rr_kernel_fn mw_kernel;         /* non-null: the CTA-per-stack kernel solves this shape (takes precedence) */
    int32_t block;                  /* threads per CTA of the chosen kernel */
    osb_options opt;
    uint32_t active_mask;
    double eps_level[OSB_MAX_LEVELS];   /* absolute diagonal shift per level */
    int32_t max_batch, device, sm_count;
    int32_t es;                     /* bytes per element: 8 (fp64) or 4 (fp32) */
    int32_t stride_e;               /* elements between instance blobs          */
    int64_t smem_bytes;
    int32_t ctas_per_sm, kernel_regs;
    char kernel_name[64];
    int64_t launches;
    char* d_blobs; char* d_x; int32_t* d_status; char* d_xlev; int32_t* d_iters; /* element type per es */
    cudaStream_t stream;            /* solve_host: first compute stream (events may be recorded on it) */
    cudaStream_t s_h2d, s_d2h, s_k2; /* solve_host pipeline: copy-in, copy-out, second compute stream  */
    cudaEvent_t ev_h2d[OSB_PIPE_CHUNKS], ev_k[OSB_PIPE_CHUNKS], ev_begin, ev_end;
    osb_pipe_slot ps[2];            /* submitted calls in flight: ps_head is the oldest, ps_count of them (at most 2) */
    int32_t ps_head, ps_count;
    /* wire format (osb_wire_configure): the dynamic runs of a blob, packed; the static rest lives in w_tmpl */
    int32_t w_ready, w_nseg, w_stride;                       /* w_stride: elements per instance on the wire */
    int32_t w_src[OSB_WIRE_MAX_SEGMENTS], w_dst[OSB_WIRE_MAX_SEGMENTS], w_cnt[OSB_WIRE_MAX_SEGMENTS];
    char* w_tmpl;                                            /* device: one blob (stride_e elements of the engine's precision) */
    /* fused all-gather (osb_gather_*): this rank's buffer [x_all | status_all | flags] and the
     * opened buffers of the peers */
    char* g_base; int64_t g_batch, g_first, g_off_status, g_off_flags, g_half;   /* two halves [x | status], by step parity */
    int32_t g_rank, g_world;
    int32_t clock_khz;               /* SM clock (barrier timeout in cycles); queried once: the attribute is slow */
    uint32_t g_epoch;
    uint32_t g_step;                 /* fused solves launched so far: step parity picks the gather half */
    char* g_peer[OSB_MAX_PEERS];     /* base of rank r's buffer in this process (own: g_base) */
    int32_t* g_fail;                 /* device word: barriers that timed out                   */
    /* kinematic tree of the Cartesian / CoM producers (device tables, depth-sorted) */
    int32_t k_joints, k_dof, k_depths;
    char* k_dev;                     /* one allocation holding every table                     */
    int32_t *k_parent, *k_type, *k_dofmap, *k_depth_start;
    double *k_axis, *k_origin, *k_mass, *k_com, *k_msub;
    /* state-input path (osb_states_configure): device copies of the constants, the task list, the q buffer */
    double* s_dev; double* s_q; int32_t s_ready, s_ntasks;
    osb_posture_inputs s_posture;
    osb_kin_task_desc s_tasks[OSB_KIN_MAX_TASKS];
    int32_t k_perm[OSB_KIN_MAX_JOINTS];      /* user joint index -> engine order                */
    int32_t k_hparent[OSB_KIN_MAX_JOINTS];   /* host copy of the sorted parent table            */
};


extern "C" void osb_engine_destroy(osb_engine* e);

extern "C" int64_t osb_smem_bytes_per_instance(const osb_stack_desc* desc)
{
    osb_layout L; sot_workspace w; rr_workspace rr; mw_workspace mw;
    if (osb_blob_layout(desc, &L) != OSB_SUCCESS) return -1;
    const int64_t rr_bytes = rr_plan_workspace(desc, &L, &rr, 8);
    if (rr_bytes >= 0 && rr.rpl == 1) return rr_bytes;
    const int64_t mw_bytes = mw_plan_workspace(desc, &L, &mw, 8);
    if (mw_bytes >= 0) return mw_bytes;
    return rr_bytes >= 0 ? rr_bytes : sot_plan_workspace(desc, &L, &w, 8);
}

extern "C" const char* osb_last_error(void) { return g_err; }
extern "C" const char* osb_version(void) { return "opensot_b200 0.1 (sm_100a)"; }
extern "C" int osb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static int engine_create(osb_engine** out, const osb_stack_desc* desc, int32_t max_batch, int32_t device, int es)
{
    if (!out || !desc || max_batch < 1) return fail(OSB_E_INVALID, "bad arguments%s", "");
    *out = nullptr;
    osb_layout L;
    int rc = osb_blob_layout(desc, &L);
    if (rc != OSB_SUCCESS) return fail(rc, "invalid stack descriptor%s", "");
    if (es == 4) {
        /* A level on the DENSE path forms Hz = (A Z)' W (A Z) + eps I explicitly. When its task has fewer
         * rows than variables, part of Hz is eps alone (1e-6 at best in fp32) under entries of order
         * |A|^2 u ~ 1e-6: pivots go negative and the eps-held directions of the answer are noise. The
         * LQ path never forms Hz and has no such limit; such stacks need the fp64 engine. */
        for (int k = 0; k < desc->n_levels; ++k)
            if (sot_level_path_of(desc, k) == SOT_PATH_DENSE && desc->task_kind[k] == OSB_TASK_DENSE &&
                desc->m[k] + desc->m_reg < desc->n)
                return fail(OSB_E_PRECISION, "fp32 engine: a dense-weight / regularised level with fewer task rows than "
                                             "variables needs the fp64 engine (explicit reduced Hessian)%s", "");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) {
        cudaGetLastError();
        return fail(OSB_E_NO_DEVICE, "no CUDA device: the engine has no CPU fallback%s", "");
    }
    CUDA_OK(cudaSetDevice(device));
    osb_engine* e = new (std::nothrow) osb_engine();
    if (!e) return fail(OSB_E_INVALID, "out of memory%s", "");
    memset(e, 0, sizeof(*e));
    e->d = *desc; e->L = L; e->max_batch = max_batch; e->device = device;
    e->es = es;
    e->stride_e = (es == 8) ? L.stride : ((L.stride + 3) & ~3);
    /* fp32: the default shift (4.4e-11) is far below single-precision resolution of the
     * 1/sqrt(eps) column scaling; use an fp32-appropriate one */
    if (es == 4 && e->d.eps < OSB_F32_MIN_EPS) e->d.eps = OSB_F32_MIN_EPS;
    for (int k = 0; k < OSB_MAX_LEVELS; ++k) e->eps_level[k] = e->d.eps;
    e->smem_bytes = sot_plan_workspace(desc, &L, &e->w, es);
    {   /* shapes inside the register-row kernel's instances are solved by it (OSB_KERNEL=generic
         * in the environment keeps the generic kernel, for A/B measurements) */
        const char* force = getenv("OSB_KERNEL");
        const bool want_generic = force && strcmp(force, "generic") == 0, want_rr = force && strcmp(force, "rr") == 0;
        e->block = 32;
        bool wh = false;        /* a dense weight on a dense task somewhere: the instances with rr_whiten */
        for (int k = 0; k < desc->n_levels; ++k)
            wh = wh || (desc->m_reg == 0 && desc->task_kind[k] == OSB_TASK_DENSE && desc->weight_kind[k] == OSB_W_DENSE);
        const int64_t rr_bytes = rr_plan_workspace(desc, &L, &e->rr, es);
        if (rr_bytes >= 0 && !want_generic) {
            e->rr_kernel = rr_pick_kernel(es, e->rr.np, e->rr.rpl, wh);
            if (e->rr_kernel) e->smem_bytes = rr_bytes;
        }
        /* more than 32 rows of y (two rows per lane on the register-row kernel) or more than 38 variables
         * (generic kernel): one CTA of 2-3 warps per stack, one row per thread (OSB_KERNEL=rr keeps the
         * two-rows-per-lane instance, for A/B measurements) */
        const int64_t mw_bytes = mw_plan_workspace(desc, &L, &e->mw, es);
        if (mw_bytes >= 0 && !want_generic && !want_rr && !(e->rr_kernel && e->rr.rpl == 1)) {
            e->mw_kernel = mw_pick_kernel(es, e->mw.np, e->mw.nw, wh);
            if (e->mw_kernel) { e->rr_kernel = nullptr; e->smem_bytes = mw_bytes; e->block = 32 * e->mw.nw; }
        }
    }
    e->opt.max_iter = 0; e->opt.store_levels = 0;
    e->opt.tol_feas = (es == 8) ? 1e-12 : 1e-6;
    e->opt.tol_rank = (es == 8) ? 1e-11 : 1e-5;
    e->active_mask = (desc->n_levels >= 32) ? 0xffffffffu : ((1u << desc->n_levels) - 1u);
    cudaDeviceProp prop;
    CUDA_OK_E(cudaGetDeviceProperties(&prop, device));
    e->sm_count = prop.multiProcessorCount;
    e->clock_khz = prop.clockRate > 0 ? prop.clockRate : 2000000;
    if (const char* pad = getenv("OSB_DEV_SMEM_PAD"))   /* development aid: throughput against resident warps (tools/occupancy_sweep.py) */
        e->smem_bytes += atoll(pad);
    if (e->smem_bytes > (int64_t)prop.sharedMemPerBlockOptin) {
        osb_engine_destroy(e);
        return fail(OSB_E_TOO_LARGE, "stack needs more shared memory per warp than one SM offers%s", "");
    }
    const void* kern = e->mw_kernel ? (const void*)e->mw_kernel : e->rr_kernel ? (const void*)e->rr_kernel
                                    : ((es == 8) ? (const void*)sot_solve_kernel : (const void*)sot_solve_kernel_f32);
    {   /* the opt-in shared-memory size is an attribute of the FUNCTION (per device): engines of different shapes share a
         * kernel, so it only ever grows (a smaller engine created later must not shrink it under a larger one) */
        static std::mutex mu;
        static std::map<std::pair<int, const void*>, int64_t> granted;
        std::lock_guard<std::mutex> lock(mu);
        int64_t& have = granted[std::make_pair(device, kern)];
        if (e->smem_bytes > have) {
            CUDA_OK_E(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
            have = e->smem_bytes;
        }
    }
    cudaFuncAttributes fa;
    CUDA_OK_E(cudaFuncGetAttributes(&fa, kern));
    e->kernel_regs = fa.numRegs;
    int occ = 0;
    CUDA_OK_E(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, e->block, (size_t)e->smem_bytes));
    e->ctas_per_sm = occ;
    const size_t B = (size_t)max_batch;
    CUDA_OK_E(cudaMalloc(&e->d_blobs, B * e->stride_e * es));
    CUDA_OK_E(cudaMalloc(&e->d_x, B * desc->n * es));
    CUDA_OK_E(cudaMalloc(&e->d_status, B * sizeof(int32_t)));
    CUDA_OK_E(cudaMalloc(&e->d_iters, B * sizeof(int32_t)));
    CUDA_OK_E(cudaMalloc(&e->d_xlev, B * desc->n_levels * desc->n * es));
    /* slices of levels that are inactive or follow a failed level are never written by the kernels */
    CUDA_OK_E(cudaMemset(e->d_xlev, 0, B * desc->n_levels * desc->n * es));
    CUDA_OK_E(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CUDA_OK_E(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking));
    CUDA_OK_E(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking));
    CUDA_OK_E(cudaStreamCreateWithFlags(&e->s_k2, cudaStreamNonBlocking));
    for (int i = 0; i < OSB_PIPE_CHUNKS; ++i) {
        CUDA_OK_E(cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDefault));
        CUDA_OK_E(cudaEventCreateWithFlags(&e->ev_k[i], cudaEventDefault));
    }
    CUDA_OK_E(cudaEventCreateWithFlags(&e->ev_begin, cudaEventDefault));
    CUDA_OK_E(cudaEventCreateWithFlags(&e->ev_end, cudaEventDefault));
    *out = e;
    return OSB_SUCCESS;
}

extern "C" int osb_engine_create(osb_engine** out, const osb_stack_desc* desc, int32_t max_batch,
                                 int32_t device)
{
    return engine_create(out, desc, max_batch, device, 8);
}
extern "C" int osb_engine_create_f32(osb_engine** out, const osb_stack_desc* desc, int32_t max_batch,
                                     int32_t device)
{
    return engine_create(out, desc, max_batch, device, 4);
}
extern "C" int32_t osb_blob_stride_f32(const osb_stack_desc* desc)
{
    osb_layout L;
    if (osb_blob_layout(desc, &L) != OSB_SUCCESS) return -1;
    return (L.stride + 3) & ~3;
}

extern "C" void osb_engine_destroy(osb_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->d_blobs); cudaFree(e->d_x); cudaFree(e->d_status); cudaFree(e->d_iters);
    cudaFree(e->d_xlev);
    for (int r = 0; r < OSB_MAX_PEERS; ++r)
        if (e->g_peer[r] && e->g_peer[r] != e->g_base) cudaIpcCloseMemHandle(e->g_peer[r]);
    if (e->k_dev) cudaFree(e->k_dev);
    if (e->s_dev) cudaFree(e->s_dev);
    if (e->s_q) cudaFree(e->s_q);
    if (e->g_base) cudaFree(e->g_base);
    if (e->g_fail) cudaFree(e->g_fail);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
    if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
    if (e->s_k2) cudaStreamDestroy(e->s_k2);
    for (int i = 0; i < OSB_PIPE_CHUNKS; ++i) {
        if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
        if (e->ev_k[i]) cudaEventDestroy(e->ev_k[i]);
    }
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    {
        osb_pipe_slot& s1 = e->ps[1];
        cudaFree(s1.d_blobs); cudaFree(s1.d_x); cudaFree(s1.d_status); cudaFree(s1.d_iters); cudaFree(s1.d_xlev); cudaFree(s1.s_q);
        cudaFree(s1.d_wire); cudaFree(e->ps[0].d_wire); cudaFree(e->w_tmpl);
        for (int i = 0; i < OSB_PIPE_CHUNKS; ++i) {
            if (s1.ev_h2d[i]) cudaEventDestroy(s1.ev_h2d[i]);
            if (s1.ev_k[i]) cudaEventDestroy(s1.ev_k[i]);
        }
        if (s1.ev_end) cudaEventDestroy(s1.ev_end);
    }
    delete e;
}

extern "C" int osb_set_options(osb_engine* e, const osb_options* opt)
{
    if (!e || !opt) return fail(OSB_E_INVALID, "bad arguments%s", "");
    e->opt = *opt;
    if (!(e->opt.tol_feas > 0)) e->opt.tol_feas = (e->es == 8) ? 1e-12 : 1e-6;
    if (!(e->opt.tol_rank > 0)) e->opt.tol_rank = (e->es == 8) ? 1e-11 : 1e-5;
    return OSB_SUCCESS;
}

extern "C" int osb_set_active_levels(osb_engine* e, const int32_t* active)
{
    if (!e || !active) return fail(OSB_E_INVALID, "bad arguments%s", "");
    uint32_t mask = 0;
    for (int k = 0; k < e->d.n_levels; ++k) if (active[k]) mask |= (1u << k);
    if (mask == 0) return fail(OSB_E_INVALID, "at least one level must stay active%s", "");
    e->active_mask = mask;
    return OSB_SUCCESS;
}

extern "C" int osb_set_eps_level(osb_engine* e, int32_t level, double eps)
{
    if (!e || level >= e->d.n_levels || !(eps >= 0.0)) return fail(OSB_E_INVALID, "bad arguments (eps must be >= 0, level < n_levels)%s", "");
    if (e->es == 4 && eps < OSB_F32_MIN_EPS) eps = OSB_F32_MIN_EPS;
    for (int k = 0; k < e->d.n_levels; ++k)
        if (level < 0 || k == level) e->eps_level[k] = eps;
    if (level < 0) e->d.eps = eps;
    return OSB_SUCCESS;
}
extern "C" double osb_get_eps_level(osb_engine* e, int32_t level)
{
    return (e && level >= 0 && level < e->d.n_levels) ? e->eps_level[level] : -1.0;
}
extern "C" int osb_get_options(osb_engine* e, osb_options* out)
{
    if (!e || !out) return fail(OSB_E_INVALID, "bad arguments%s", "");
    *out = e->opt;
    return OSB_SUCCESS;
}

/* page-locked host memory for blobs and solutions: the copy engine reads pageable memory through a
 * staging buffer at a fraction of the PCIe rate */
extern "C" void* osb_host_alloc(int64_t bytes)
{
    void* ptr = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&ptr, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return ptr;
}
extern "C" void osb_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

static int launch_solve(osb_engine* e, const void* d_blobs, int32_t batch, void* d_x,
                        int32_t* d_status, void* d_xlev, int32_t* d_iters, cudaStream_t st, bool to_peers = false)
{
    sot_params p;
    memset(&p, 0, sizeof(p));
    if (to_peers) {
        p.peer_first = e->g_first;
        const int64_t half = (int64_t)(e->g_step & 1u) * e->g_half;       /* this step's half */
        for (int r = 0; r < e->g_world; ++r) {
            if (r == e->g_rank) continue;
            p.peer_x[p.n_peers] = e->g_peer[r] + half;
            p.peer_status[p.n_peers] = (int32_t*)(e->g_peer[r] + half + e->g_off_status);
            ++p.n_peers;
        }
    }
    p.d = e->d; p.L = e->L; p.w = e->w; p.rr = e->rr; p.mw = e->mw;
    p.max_iter = e->opt.max_iter; p.store_levels = (e->opt.store_levels && d_xlev) ? 1 : 0;
    p.tol_feas = e->opt.tol_feas; p.tol_rank = e->opt.tol_rank;
    for (int k = 0; k < OSB_MAX_LEVELS; ++k) p.eps_level[k] = e->eps_level[k];
    p.active_mask = e->active_mask; p.batch = batch;
    p.blobs = d_blobs; p.x = d_x; p.status = d_status; p.x_levels = d_xlev; p.iters = d_iters;
    if (p.store_levels)   /* levels that are inactive or follow a failed level are not written by the kernels: zeros */
        CUDA_OK(cudaMemsetAsync(d_xlev, 0, (size_t)batch * e->d.n_levels * e->d.n * e->es, st));
    if (e->mw_kernel) e->mw_kernel<<<batch, e->block, (size_t)e->smem_bytes, st>>>(p);
    else if (e->rr_kernel) e->rr_kernel<<<batch, 32, (size_t)e->smem_bytes, st>>>(p);
    else if (e->es == 8) sot_solve_kernel<<<batch, 32, (size_t)e->smem_bytes, st>>>(p);
    else sot_solve_kernel_f32<<<batch, 32, (size_t)e->smem_bytes, st>>>(p);
    CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return OSB_SUCCESS;
}

static int solve_device_impl(osb_engine* e, const void* d_blobs, int32_t batch, void* d_x,
                             int32_t* d_status, void* d_x_levels, int32_t* d_iters,
                             void* cuda_stream, int es)
{
    if (!e || !d_blobs || !d_x || !d_status || batch < 1)
        return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->es != es) return fail(OSB_E_INVALID, "engine precision does not match this entry point%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    return launch_solve(e, d_blobs, batch, d_x, d_status, d_x_levels, d_iters, (cudaStream_t)cuda_stream);
}

/* ---- constant parts of the blobs: templates and the wire format ---------------------------------------- */
__global__ void sot_template_kernel(double* blobs, const double* tmpl, int stride, long long total)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) blobs[i] = tmpl[i % stride];
}
__global__ void sot_template_kernel_f32(float* blobs, const float* tmpl, int stride, long long total)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) blobs[i] = tmpl[i % stride];
}
/* wire -> blobs: the dynamic runs of every instance go to their places (thread per wire element; reads are coalesced,
 * writes are runs of hundreds of elements); the static rest of the blobs was written once from the template */
struct wire_params {
    int32_t nseg, wire_stride, blob_stride;
    int32_t src[OSB_WIRE_MAX_SEGMENTS], dst[OSB_WIRE_MAX_SEGMENTS], cnt[OSB_WIRE_MAX_SEGMENTS];
    long long total;
};
template <typename T>
__global__ void sot_wire_expand_kernel(const T* __restrict__ wire, T* __restrict__ blobs, const __grid_constant__ wire_params g)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.total) return;
    const long long inst = idx / g.wire_stride;
    const int k = (int)(idx - inst * g.wire_stride);
#pragma unroll 1
    for (int s = 0; s < g.nseg; ++s)
        if (k >= g.src[s] && k < g.src[s] + g.cnt[s]) { blobs[inst * g.blob_stride + g.dst[s] + (k - g.src[s])] = wire[idx]; break; }
}

/* makes the constant part of a slot's blobs what `tag` needs (no-op when it already is); enqueued on `st` */
static int ensure_template(osb_engine* e, osb_pipe_slot& ps, int tag, cudaStream_t st)
{
    if (ps.tag == tag) return OSB_SUCCESS;
    const long long cells = (long long)e->max_batch * e->stride_e;
    const unsigned grid = (unsigned)((cells + 255) / 256);
    if (tag == OSB_TAG_STATES)
        sot_template_kernel<<<grid, 256, 0, st>>>((double*)ps.d_blobs, e->s_dev, e->stride_e, cells);
    else if (e->es == 8)
        sot_template_kernel<<<grid, 256, 0, st>>>((double*)ps.d_blobs, (const double*)e->w_tmpl, e->stride_e, cells);
    else
        sot_template_kernel_f32<<<grid, 256, 0, st>>>((float*)ps.d_blobs, (const float*)e->w_tmpl, e->stride_e, cells);
    CUDA_OK(cudaGetLastError());
    ps.tag = tag;
    return OSB_SUCCESS;
}

/* the buffers of pipeline slot `which`: slot 0 is the engine's own set, slot 1 is allocated on first use */
static int pipe_slot(osb_engine* e, int which, osb_pipe_slot** out)
{
    osb_pipe_slot& s = e->ps[which];
    if (which == 0) {
        s.d_blobs = e->d_blobs; s.d_x = e->d_x; s.d_status = e->d_status; s.d_xlev = e->d_xlev; s.d_iters = e->d_iters; s.s_q = e->s_q;   /* (tag and d_wire persist) */
        for (int i = 0; i < OSB_PIPE_CHUNKS; ++i) { s.ev_h2d[i] = e->ev_h2d[i]; s.ev_k[i] = e->ev_k[i]; }
        s.ev_end = e->ev_end; s.ready = 1;
    } else {
        const size_t B = (size_t)e->max_batch, es = (size_t)e->es;
        if (!s.ready) {
            /* all or nothing: a failure half-way gives back what was taken (the next submit tries again) */
            bool ok = cudaMalloc(&s.d_blobs, B * e->stride_e * es) == cudaSuccess
                   && cudaMalloc(&s.d_x, B * e->d.n * es) == cudaSuccess
                   && cudaMalloc(&s.d_status, B * sizeof(int32_t)) == cudaSuccess
                   && cudaMalloc(&s.d_iters, B * sizeof(int32_t)) == cudaSuccess
                   && cudaMalloc(&s.d_xlev, B * e->d.n_levels * e->d.n * es) == cudaSuccess
                   && cudaMemset(s.d_xlev, 0, B * e->d.n_levels * e->d.n * es) == cudaSuccess
                   && cudaEventCreateWithFlags(&s.ev_end, cudaEventDefault) == cudaSuccess;
            for (int i = 0; ok && i < OSB_PIPE_CHUNKS; ++i)
                ok = cudaEventCreateWithFlags(&s.ev_h2d[i], cudaEventDefault) == cudaSuccess
                  && cudaEventCreateWithFlags(&s.ev_k[i], cudaEventDefault) == cudaSuccess;
            if (!ok) {
                const cudaError_t err = cudaGetLastError();
                cudaFree(s.d_blobs); cudaFree(s.d_x); cudaFree(s.d_status); cudaFree(s.d_iters); cudaFree(s.d_xlev);
                for (int i = 0; i < OSB_PIPE_CHUNKS; ++i) {
                    if (s.ev_h2d[i]) cudaEventDestroy(s.ev_h2d[i]);
                    if (s.ev_k[i]) cudaEventDestroy(s.ev_k[i]);
                }
                if (s.ev_end) cudaEventDestroy(s.ev_end);
                memset(&s, 0, sizeof(s));
                return fail(OSB_E_CUDA, "second set of device buffers (two calls in flight): %s", cudaGetErrorString(err));
            }
            s.ready = 1;
        }
        if (e->s_ready && !s.s_q) CUDA_OK(cudaMalloc(&s.s_q, B * e->d.n * sizeof(double)));
    }
    if (e->w_ready && !s.d_wire) CUDA_OK(cudaMalloc(&s.d_wire, (size_t)e->max_batch * e->w_stride * e->es));
    *out = &s;
    return OSB_SUCCESS;
}

/*
 * Host-buffer entry, enqueue part. The batch is cut into chunks and pipelined over three engines of the
 * GPU: chunk c+1 crosses PCIe (copy stream) while chunk c is being solved (two compute
 * streams, so that the tail of one chunk's grid overlaps the head of the next) and the
 * solutions of chunk c-1 travel back (copy-out stream). `fence` (the blocking entry points): everything
 * is ordered behind e->stream, so a caller that records events on osb_engine_stream() brackets the whole
 * call. Without it (osb_solve_host_submit) the copy-in of this call follows the copy-in of the previous
 * one directly: its blobs cross PCIe while the previous call's last chunks are still being solved.
 */
static int host_enqueue(osb_engine* e, osb_pipe_slot& ps, const char* blobs, int32_t batch, char* x,
                        int32_t* status, char* x_levels, bool fence, size_t* bound_out, int* chunks_out, bool wire = false)
{
    const int es = e->es;
    /* byte strides per instance */
    const size_t stride = (size_t)e->stride_e * es, n = (size_t)e->d.n * es, nl = (size_t)e->d.n_levels;
    const size_t wstride = (size_t)e->w_stride * es;
    if (wire) {     /* the static part of this slot's blobs, once (on the copy stream: the kernels wait for its events) */
        int rc = ensure_template(e, ps, OSB_TAG_WIRE, e->s_h2d);
        if (rc != OSB_SUCCESS) return rc;
    } else ps.tag = OSB_TAG_NONE;
    /* chunk boundaries: halves, down to OSB_PIPE_LAST_CHUNK instances. The copy engine is the
     * bottleneck (the solve of a chunk ends long before the next chunk has arrived), so the call
     * ends one solve after the last byte has crossed PCIe: that last solve should be as short
     * as a solve gets (one stack is ~50 us whatever the batch), while the first chunks should be
     * large (fewer launches and events). */
    size_t bound[OSB_PIPE_CHUNKS + 1];
    int chunks = 0;
    if (!fence) {
        /* submitted calls overlap ACROSS calls (the next call's blobs cross PCIe while this one is solved), so a call
         * is one copy-in, one launch, one copy-out: the copy engine runs back to back at the rate of large transfers
         * (54.9 GB/s measured against 48 GB/s for the five chunks of the blocking call) */
        static const int want = getenv("OSB_SUBMIT_CHUNKS") ? atoi(getenv("OSB_SUBMIT_CHUNKS")) : 1;
        const int k = want < 1 ? 1 : (want > OSB_PIPE_CHUNKS ? OSB_PIPE_CHUNKS : want);
        for (int c = 0; c <= k; ++c) bound[c] = (size_t)batch * c / k;
        chunks = k;
        if ((size_t)batch < (size_t)k) { bound[1] = (size_t)batch; chunks = 1; }
    } else if (const char* pat = getenv("OSB_PIPE_PATTERN")) {   /* development aid: "2048,1024,512,512" (must sum to the batch) */
        size_t at = 0; bound[0] = 0;
        while (*pat && chunks < OSB_PIPE_CHUNKS) {
            at += (size_t)strtol(pat, (char**)&pat, 10);
            bound[++chunks] = at;
            if (*pat == ',') ++pat;
        }
        if (at != (size_t)batch) { bound[1] = (size_t)batch; chunks = 1; }
    } else if (batch >= 2048 && batch <= 8192) {
        /* measured at 4096 stacks (tools/pipe_pattern_sweep.py): 5/8, 2/8, 1/8 of the batch ends 4 % earlier than five
         * halving chunks (0.488 against 0.509 ms): every chunk boundary is a gap on the copy engine */
        bound[0] = 0; bound[1] = ((size_t)batch * 5 / 8) & ~(size_t)31; bound[2] = ((size_t)batch * 7 / 8) & ~(size_t)31; bound[3] = (size_t)batch;
        chunks = 3;
    } else {
        size_t at = 0;
        bound[0] = 0;
        while (at < (size_t)batch && chunks < OSB_PIPE_CHUNKS) {
            size_t left = (size_t)batch - at, take = (left + 1) / 2;
            if (take < OSB_PIPE_LAST_CHUNK || chunks == OSB_PIPE_CHUNKS - 1) take = left;
            if (left - take < OSB_PIPE_LAST_CHUNK / 2) take = left;
            at += take;
            bound[++chunks] = at;
        }
    }
    const bool levels = x_levels && e->opt.store_levels;
    if (fence) {
        CUDA_OK(cudaEventRecord(e->ev_begin, e->stream));
        CUDA_OK(cudaStreamWaitEvent(e->s_h2d, e->ev_begin, 0));
        CUDA_OK(cudaStreamWaitEvent(e->s_k2, e->ev_begin, 0));
        CUDA_OK(cudaStreamWaitEvent(e->s_d2h, e->ev_begin, 0));
    }
    for (int c = 0; c < chunks; ++c) {
        const size_t lo = bound[c], hi = bound[c + 1], cnt = hi - lo;
        /* the two slots start on different compute streams: the head of a submitted call overlaps the tail of the previous one */
        cudaStream_t sk = ((c + (&ps == &e->ps[1] ? 1 : 0)) & 1) ? e->s_k2 : e->stream;
        if (wire)
            CUDA_OK(cudaMemcpyAsync(ps.d_wire + lo * wstride, blobs + lo * wstride, cnt * wstride, cudaMemcpyHostToDevice, e->s_h2d));
        else
            CUDA_OK(cudaMemcpyAsync(ps.d_blobs + lo * stride, blobs + lo * stride, cnt * stride,
                                    cudaMemcpyHostToDevice, e->s_h2d));
        CUDA_OK(cudaEventRecord(ps.ev_h2d[c], e->s_h2d));
        CUDA_OK(cudaStreamWaitEvent(sk, ps.ev_h2d[c], 0));
        if (wire) {
            wire_params g;
            g.nseg = e->w_nseg; g.wire_stride = e->w_stride; g.blob_stride = e->stride_e; g.total = (long long)cnt * e->w_stride;
            for (int q = 0; q < OSB_WIRE_MAX_SEGMENTS; ++q) { g.src[q] = e->w_src[q]; g.dst[q] = e->w_dst[q]; g.cnt[q] = e->w_cnt[q]; }
            const unsigned grid = (unsigned)((g.total + 255) / 256);
            if (es == 8) sot_wire_expand_kernel<double><<<grid, 256, 0, sk>>>((const double*)(ps.d_wire + lo * wstride), (double*)(ps.d_blobs + lo * stride), g);
            else sot_wire_expand_kernel<float><<<grid, 256, 0, sk>>>((const float*)(ps.d_wire + lo * wstride), (float*)(ps.d_blobs + lo * stride), g);
            CUDA_OK(cudaGetLastError());
            e->launches += 1;
        }
        int rc = launch_solve(e, ps.d_blobs + lo * stride, (int32_t)cnt, ps.d_x + lo * n, ps.d_status + lo,
                              levels ? ps.d_xlev + lo * nl * n : nullptr, ps.d_iters + lo, sk);
        if (rc != OSB_SUCCESS) return rc;
        CUDA_OK(cudaEventRecord(ps.ev_k[c], sk));
        CUDA_OK(cudaStreamWaitEvent(e->s_d2h, ps.ev_k[c], 0));
        CUDA_OK(cudaMemcpyAsync(x + lo * n, ps.d_x + lo * n, cnt * n, cudaMemcpyDeviceToHost, e->s_d2h));
        CUDA_OK(cudaMemcpyAsync(status + lo, ps.d_status + lo, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_d2h));
        if (levels)
            CUDA_OK(cudaMemcpyAsync(x_levels + lo * nl * n, ps.d_xlev + lo * nl * n, cnt * nl * n,
                                    cudaMemcpyDeviceToHost, e->s_d2h));
    }
    CUDA_OK(cudaEventRecord(ps.ev_end, e->s_d2h));
    if (bound_out) for (int c = 0; c <= chunks; ++c) bound_out[c] = bound[c];
    if (chunks_out) *chunks_out = chunks;
    return OSB_SUCCESS;
}

static int solve_host_impl(osb_engine* e, const void* blobs_v, int32_t batch, void* x_v,
                           int32_t* status, void* xl_v, int es)
{
    if (!e || !blobs_v || !x_v || !status || batch < 1 || batch > e->max_batch)
        return fail(OSB_E_INVALID, "bad arguments (batch must be 1..max_batch)%s", "");
    if (e->es != es) return fail(OSB_E_INVALID, "engine precision does not match this entry point%s", "");
    if (e->ps_count > 0) return fail(OSB_E_INVALID, "submitted calls are still in flight: osb_solve_wait first%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    osb_pipe_slot* ps = nullptr;
    int rc = pipe_slot(e, 0, &ps);
    if (rc != OSB_SUCCESS) return rc;
    size_t bound[OSB_PIPE_CHUNKS + 1]; int chunks = 0;
    rc = host_enqueue(e, *ps, (const char*)blobs_v, batch, (char*)x_v, status, (char*)xl_v, true, bound, &chunks);
    if (rc != OSB_SUCCESS) return rc;
    CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_end, 0));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    if (getenv("OSB_PIPE_TRACE")) {   /* development aid: when did each chunk arrive / finish (ms after the call began) */
        for (int c = 0; c < chunks; ++c) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, e->ev_begin, e->ev_h2d[c]); cudaEventElapsedTime(&b, e->ev_begin, e->ev_k[c]);
            fprintf(stderr, "chunk %d [%zu,%zu): h2d done %.3f solved %.3f\n", c, bound[c], bound[c + 1], a, b);
        }
        float t = 0; cudaEventElapsedTime(&t, e->ev_begin, e->ev_end);
        fprintf(stderr, "d2h done %.3f\n", t);
    }
    return OSB_SUCCESS;
}

/* ---- two host-buffer calls in flight: the blobs of call k+1 cross PCIe while call k is solved ---------- */
static int submit_slot(osb_engine* e, osb_pipe_slot** ps)
{
    if (e->ps_count >= 2) return fail(OSB_E_INVALID, "two submitted calls are already in flight: osb_solve_wait first%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    return pipe_slot(e, (e->ps_head + e->ps_count) & 1, ps);
}
static int solve_host_submit_impl(osb_engine* e, const void* blobs_v, int32_t batch, void* x_v, int32_t* status, void* xl_v, int es)
{
    if (!e || !blobs_v || !x_v || !status || batch < 1 || batch > e->max_batch)
        return fail(OSB_E_INVALID, "bad arguments (batch must be 1..max_batch)%s", "");
    if (e->es != es) return fail(OSB_E_INVALID, "engine precision does not match this entry point%s", "");
    osb_pipe_slot* ps = nullptr;
    int rc = submit_slot(e, &ps);
    if (rc != OSB_SUCCESS) return rc;
    rc = host_enqueue(e, *ps, (const char*)blobs_v, batch, (char*)x_v, status, (char*)xl_v, false, nullptr, nullptr);
    if (rc != OSB_SUCCESS) return rc;
    e->ps_count += 1;
    return OSB_SUCCESS;
}
extern "C" int osb_solve_host_submit(osb_engine* e, const double* blobs, int32_t batch, double* x, int32_t* status, double* x_levels)
{
    return solve_host_submit_impl(e, blobs, batch, x, status, x_levels, 8);
}
extern "C" int osb_solve_host_submit_f32(osb_engine* e, const float* blobs, int32_t batch, float* x, int32_t* status, float* x_levels)
{
    return solve_host_submit_impl(e, blobs, batch, x, status, x_levels, 4);
}
/* ---- wire format: only the fields that change every cycle cross PCIe ----------------------------------- */
extern "C" int osb_wire_configure(osb_engine* e, const osb_wire_segment* static_segments, int32_t n_static, const double* template_blob)
{
    if (!e || n_static < 0 || (n_static > 0 && !static_segments) || !template_blob) return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->ps_count > 0) return fail(OSB_E_INVALID, "submitted calls are still in flight: osb_solve_wait first%s", "");
    const int32_t stride = e->L.stride;
    /* dynamic runs = the complement of the static segments (sorted, disjoint, inside the blob) */
    int32_t nseg = 0, at = 0, wire = 0;
    int32_t src[OSB_WIRE_MAX_SEGMENTS], dst[OSB_WIRE_MAX_SEGMENTS], cnt[OSB_WIRE_MAX_SEGMENTS];
    for (int i = 0; i <= n_static; ++i) {
        const int32_t lo = (i < n_static) ? static_segments[i].offset : stride;
        const int32_t len = (i < n_static) ? static_segments[i].count : 0;
        if (lo < at || len < 0 || lo + len > stride) return fail(OSB_E_INVALID, "static segments must be sorted, disjoint and inside the blob%s", "");
        if (lo > at) {
            if (nseg == OSB_WIRE_MAX_SEGMENTS) return fail(OSB_E_TOO_LARGE, "more than 16 dynamic runs%s", "");
            src[nseg] = wire; dst[nseg] = at; cnt[nseg] = lo - at; wire += lo - at; ++nseg;
        }
        at = lo + len;
    }
    if (nseg == 0) return fail(OSB_E_INVALID, "nothing left to send: every field is static%s", "");
    const int32_t align = (e->es == 8) ? 2 : 4;            /* 16-byte rows on the wire */
    const int32_t wstride = (wire + align - 1) / align * align;
    CUDA_OK(cudaSetDevice(e->device));
    if (!e->w_tmpl) CUDA_OK(cudaMalloc(&e->w_tmpl, (size_t)e->stride_e * e->es));
    if (e->es == 8) {
        CUDA_OK(cudaMemcpy(e->w_tmpl, template_blob, sizeof(double) * stride, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> t32((size_t)e->stride_e, 0.0f);
        for (int i = 0; i < stride; ++i) t32[i] = (float)template_blob[i];
        CUDA_OK(cudaMemcpy(e->w_tmpl, t32.data(), sizeof(float) * e->stride_e, cudaMemcpyHostToDevice));
    }
    if (wstride != e->w_stride)          /* wire buffers are sized for the stride: give the old ones back */
        for (int q = 0; q < 2; ++q) { cudaFree(e->ps[q].d_wire); e->ps[q].d_wire = nullptr; }
    e->w_nseg = nseg; e->w_stride = wstride;
    for (int q = 0; q < OSB_WIRE_MAX_SEGMENTS; ++q) { e->w_src[q] = q < nseg ? src[q] : 0; e->w_dst[q] = q < nseg ? dst[q] : 0; e->w_cnt[q] = q < nseg ? cnt[q] : 0; }
    for (int q = 0; q < 2; ++q) if (e->ps[q].tag == OSB_TAG_WIRE) e->ps[q].tag = OSB_TAG_NONE;   /* new template */
    e->w_ready = 1;
    return OSB_SUCCESS;
}
extern "C" int32_t osb_wire_stride(osb_engine* e) { return (e && e->w_ready) ? e->w_stride : 0; }

/* host helper: full blobs (the engine's precision and stride) -> wire rows */
extern "C" int osb_wire_pack(osb_engine* e, const void* blobs, int32_t batch, void* wire)
{
    if (!e || !blobs || !wire || batch < 1) return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (!e->w_ready) return fail(OSB_E_INVALID, "call osb_wire_configure first%s", "");
    const size_t es = (size_t)e->es, bs = (size_t)e->stride_e * es, ws = (size_t)e->w_stride * es;
    for (int32_t i = 0; i < batch; ++i) {
        const char* b = (const char*)blobs + (size_t)i * bs; char* w = (char*)wire + (size_t)i * ws;
        int32_t used = 0;
        for (int q = 0; q < e->w_nseg; ++q) { memcpy(w + (size_t)e->w_src[q] * es, b + (size_t)e->w_dst[q] * es, (size_t)e->w_cnt[q] * es); used = e->w_src[q] + e->w_cnt[q]; }
        if (used < e->w_stride) memset(w + (size_t)used * es, 0, (size_t)(e->w_stride - used) * es);
    }
    return OSB_SUCCESS;
}

static int wire_call(osb_engine* e, const void* wire, int32_t batch, void* x, int32_t* status, void* x_levels, bool submit)
{
    if (!e || !wire || !x || !status || batch < 1 || batch > e->max_batch)
        return fail(OSB_E_INVALID, "bad arguments (batch must be 1..max_batch)%s", "");
    if (!e->w_ready) return fail(OSB_E_INVALID, "call osb_wire_configure first%s", "");
    osb_pipe_slot* ps = nullptr;
    int rc;
    if (submit) rc = submit_slot(e, &ps);
    else {
        if (e->ps_count > 0) return fail(OSB_E_INVALID, "submitted calls are still in flight: osb_solve_wait first%s", "");
        CUDA_OK(cudaSetDevice(e->device));
        rc = pipe_slot(e, 0, &ps);
    }
    if (rc != OSB_SUCCESS) return rc;
    rc = host_enqueue(e, *ps, (const char*)wire, batch, (char*)x, status, (char*)x_levels, !submit, nullptr, nullptr, true);
    if (rc != OSB_SUCCESS) return rc;
    if (submit) { e->ps_count += 1; return OSB_SUCCESS; }
    CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_end, 0));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return OSB_SUCCESS;
}
extern "C" int osb_solve_host_wire(osb_engine* e, const void* wire, int32_t batch, void* x, int32_t* status, void* x_levels)
{
    return wire_call(e, wire, batch, x, status, x_levels, false);
}
extern "C" int osb_solve_host_wire_submit(osb_engine* e, const void* wire, int32_t batch, void* x, int32_t* status, void* x_levels)
{
    return wire_call(e, wire, batch, x, status, x_levels, true);
}

extern "C" int osb_solve_wait(osb_engine* e)
{
    if (!e) return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->ps_count < 1) return fail(OSB_E_INVALID, "no submitted call in flight%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    osb_pipe_slot& ps = e->ps[e->ps_head];
    e->ps_head ^= 1; e->ps_count -= 1;        /* the slot is given back even if the wait reports a CUDA error */
    CUDA_OK(cudaEventSynchronize(ps.ev_end));
    return OSB_SUCCESS;
}
extern "C" int32_t osb_solve_in_flight(osb_engine* e) { return e ? e->ps_count : 0; }

extern "C" int osb_solve_device(osb_engine* e, const double* d_blobs, int32_t batch, double* d_x,
                                int32_t* d_status, double* d_x_levels, int32_t* d_iters, void* cuda_stream)
{
    return solve_device_impl(e, d_blobs, batch, d_x, d_status, d_x_levels, d_iters, cuda_stream, 8);
}
extern "C" int osb_solve_device_f32(osb_engine* e, const float* d_blobs, int32_t batch, float* d_x,
                                    int32_t* d_status, float* d_x_levels, int32_t* d_iters, void* cuda_stream)
{
    return solve_device_impl(e, d_blobs, batch, d_x, d_status, d_x_levels, d_iters, cuda_stream, 4);
}
extern "C" int osb_solve_host(osb_engine* e, const double* blobs, int32_t batch, double* x,
                              int32_t* status, double* x_levels)
{
    return solve_host_impl(e, blobs, batch, x, status, x_levels, 8);
}
extern "C" int osb_solve_host_f32(osb_engine* e, const float* blobs, int32_t batch, float* x,
                                  int32_t* status, float* x_levels)
{
    return solve_host_impl(e, blobs, batch, x, status, x_levels, 4);
}

/* ---- solve fused with the all-gather of the solutions over peer memory ---------------------- */
static size_t round256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" int osb_gather_create(osb_engine* e, int64_t global_batch, int32_t rank, int32_t world,
                                 int64_t first_instance, uint8_t handle_out[OSB_IPC_HANDLE_BYTES])
{
    if (!e || !handle_out || world < 1 || world > OSB_MAX_PEERS || rank < 0 || rank >= world ||
        global_batch < 1 || first_instance < 0 || first_instance >= global_batch)
        return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->g_base) return fail(OSB_E_INVALID, "gather buffer exists already%s", "");
    static_assert(sizeof(cudaIpcMemHandle_t) == OSB_IPC_HANDLE_BYTES, "IPC handle size");
    CUDA_OK(cudaSetDevice(e->device));
    /* TWO [x | status] halves, used by step parity: the peer stores of step s+1 go to the other half, so a rank that
     * is still reading the gathered batch of step s is never overwritten by a faster peer */
    const size_t xb = round256((size_t)global_batch * e->d.n * e->es), sb = round256((size_t)global_batch * 4);
    CUDA_OK(cudaMalloc(&e->g_base, 2 * (xb + sb) + 256));
    CUDA_OK(cudaMemset(e->g_base, 0, 2 * (xb + sb) + 256));
    CUDA_OK(cudaMalloc(&e->g_fail, sizeof(int32_t)));
    CUDA_OK(cudaMemset(e->g_fail, 0, sizeof(int32_t)));
    CUDA_OK(cudaDeviceSynchronize());
    e->g_batch = global_batch; e->g_first = first_instance; e->g_rank = rank; e->g_world = world;
    e->g_off_status = (int64_t)xb; e->g_half = (int64_t)(xb + sb); e->g_off_flags = (int64_t)(2 * (xb + sb)); e->g_epoch = 0; e->g_step = 0;
    e->g_peer[rank] = e->g_base;
    cudaIpcMemHandle_t h;
    CUDA_OK(cudaIpcGetMemHandle(&h, e->g_base));
    memcpy(handle_out, &h, sizeof(h));
    return OSB_SUCCESS;
}

extern "C" int osb_gather_connect(osb_engine* e, int32_t peer, const uint8_t handle[OSB_IPC_HANDLE_BYTES])
{
    if (!e || !handle || !e->g_base || peer < 0 || peer >= e->g_world || peer == e->g_rank)
        return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->g_peer[peer]) return fail(OSB_E_INVALID, "peer connected already%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* ptr = nullptr;
    CUDA_OK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    e->g_peer[peer] = (char*)ptr;
    return OSB_SUCCESS;
}

static int gather_ready(osb_engine* e)
{
    if (!e || !e->g_base) return fail(OSB_E_INVALID, "no gather buffer: call osb_gather_create first%s", "");
    for (int r = 0; r < e->g_world; ++r)
        if (!e->g_peer[r]) return fail(OSB_E_INVALID, "a peer is not connected: call osb_gather_connect for every other rank%s", "");
    return OSB_SUCCESS;
}

extern "C" int osb_solve_device_gather(osb_engine* e, const void* d_blobs, int32_t batch, void* cuda_stream)
{
    int rc = gather_ready(e);
    if (rc != OSB_SUCCESS) return rc;
    if (!d_blobs || batch < 1 || e->g_first + batch > e->g_batch)
        return fail(OSB_E_INVALID, "bad arguments (the shard must fit the global batch)%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    const int64_t half = (int64_t)(e->g_step & 1u) * e->g_half;
    char* own_x = e->g_base + half + (size_t)e->g_first * e->d.n * e->es;
    int32_t* own_s = (int32_t*)(e->g_base + half + e->g_off_status) + e->g_first;
    const int rc2 = launch_solve(e, d_blobs, batch, own_x, own_s, nullptr, nullptr, (cudaStream_t)cuda_stream, true);
    if (rc2 == OSB_SUCCESS) e->g_step += 1;
    return rc2;
}

extern "C" int osb_gather_barrier(osb_engine* e, void* cuda_stream, int32_t timeout_ms)
{
    int rc = gather_ready(e);
    if (rc != OSB_SUCCESS) return rc;
    CUDA_OK(cudaSetDevice(e->device));
    gather_barrier_params g;
    memset(&g, 0, sizeof(g));
    g.rank = e->g_rank; g.world = e->g_world; g.epoch = ++e->g_epoch;
    for (int r = 0; r < e->g_world; ++r) g.flags[r] = (uint32_t*)(e->g_peer[r] + e->g_off_flags);
    g.timeout_cycles = (long long)e->clock_khz * (long long)(timeout_ms > 0 ? timeout_ms : 2000);
    g.fail = e->g_fail;
    sot_gather_barrier_kernel<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(g);
    CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return OSB_SUCCESS;
}

extern "C" int osb_gather_buffers(osb_engine* e, void** d_x_all, int32_t** d_status_all)
{
    if (!e || !e->g_base) return fail(OSB_E_INVALID, "no gather buffer%s", "");
    /* the half the last closed step (or, before any barrier, the first step) gathers into */
    const int64_t half = (int64_t)((e->g_step > 0 ? e->g_step - 1 : 0) & 1u) * e->g_half;
    if (d_x_all) *d_x_all = e->g_base + half;
    if (d_status_all) *d_status_all = (int32_t*)(e->g_base + half + e->g_off_status);
    return OSB_SUCCESS;
}

extern "C" int osb_gather_barrier_failures(osb_engine* e)
{
    if (!e || !e->g_fail) return 0;
    int32_t v = 0;
    cudaSetDevice(e->device);
    if (cudaMemcpy(&v, e->g_fail, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
}

extern "C" int osb_stack_gather_device(osb_engine* e, const osb_gather_src* srcs, int32_t n_srcs,
                                       double* d_blobs, int32_t batch, void* cuda_stream)
{
    if (!e || !srcs || n_srcs < 1 || !d_blobs || batch < 1)
        return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->es != 8) return fail(OSB_E_INVALID, "device-side piling is fp64 only%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    for (int32_t first = 0; first < n_srcs; first += 16) {
        gather_params g;
        memset(&g, 0, sizeof(g));
        g.n_srcs = (n_srcs - first < 16) ? (n_srcs - first) : 16;
        for (int s = 0; s < g.n_srcs; ++s) {
            g.src[s] = srcs[first + s];
            if (g.src[s].dst_off < 0 || g.src[s].dst_off + g.src[s].rows * g.src[s].cols > e->L.stride)
                return fail(OSB_E_INVALID, "gather source writes outside the blob%s", "");
        }
        g.stride = e->L.stride; g.batch = batch; g.blobs = d_blobs;
        const unsigned grid = (unsigned)(batch < e->sm_count * 16 ? batch : e->sm_count * 16);
        sot_gather_kernel<<<grid, 128, 0, (cudaStream_t)cuda_stream>>>(g);
        CUDA_OK(cudaGetLastError());
        e->launches += 1;
    }
    return OSB_SUCCESS;
}

extern "C" int osb_produce_posture_device(osb_engine* e, const osb_posture_inputs* in, double* d_blobs,
                                          int32_t batch, void* cuda_stream)
{
    if (!e || !in || !d_blobs || batch < 1 || !in->q) return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->es != 8) return fail(OSB_E_INVALID, "joint-state producers are fp64 only%s", "");
    posture_params g;
    memset(&g, 0, sizeof(g));
    g.in = *in; g.n = e->d.n; g.stride = e->L.stride; g.batch = batch; g.blobs = d_blobs;
    g.off_b = g.off_jl = g.off_ju = g.off_vl = g.off_vu = -1;
    if (in->postural_level >= 0) {
        const int k = in->postural_level;
        if (k >= e->d.n_levels || e->d.task_kind[k] != OSB_TASK_IDENTITY || !in->q_ref)
            return fail(OSB_E_INVALID, "postural_level must name an OSB_TASK_IDENTITY level and q_ref must be given%s", "");
        g.off_b = e->L.b[k];
    }
    if (in->joint_limit_set >= 0) {
        if (in->joint_limit_set >= e->d.n_bound_sets || !in->q_min || !in->q_max)
            return fail(OSB_E_INVALID, "joint_limit_set out of range or limits missing%s", "");
        g.off_jl = e->L.l[in->joint_limit_set]; g.off_ju = e->L.u[in->joint_limit_set];
    }
    if (in->velocity_limit_set >= 0) {
        if (in->velocity_limit_set >= e->d.n_bound_sets || !in->qdot_max)
            return fail(OSB_E_INVALID, "velocity_limit_set out of range or qdot_max missing%s", "");
        g.off_vl = e->L.l[in->velocity_limit_set]; g.off_vu = e->L.u[in->velocity_limit_set];
    }
    CUDA_OK(cudaSetDevice(e->device));
    const int64_t total = (int64_t)batch * g.n;
    sot_posture_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(g);
    CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return OSB_SUCCESS;
}

/* ---- Cartesian / CoM producers ---------------------------------------------------------------- */
extern "C" int osb_kinematics_create(osb_engine* e, const osb_kin_tree* t)
{
    if (!e || !t || t->n_joints < 1 || t->n_joints > OSB_KIN_MAX_JOINTS || !t->parent || !t->type || !t->dof ||
        !t->axis || !t->origin || !t->mass || !t->com)
        return fail(OSB_E_INVALID, "bad kinematic tree (1..128 joints, every table given)%s", "");
    if (t->n_dof != e->d.n) return fail(OSB_E_INVALID, "the tree's degrees of freedom must equal the stack's variables%s", "");
    if (e->es != 8) return fail(OSB_E_INVALID, "kinematic producers are fp64 only%s", "");
    const int J = t->n_joints;
    std::vector<int> depth(J, -1);
    int maxd = 0;
    for (int j = 0; j < J; ++j) {
        const int p = t->parent[j];
        if (p >= j) return fail(OSB_E_INVALID, "kinematic tree: parents must come before their children%s", "");
        if (t->type[j] != OSB_KIN_FIXED && (t->dof[j] < 0 || t->dof[j] >= t->n_dof)) return fail(OSB_E_INVALID, "kinematic tree: dof index out of range%s", "");
        depth[j] = p < 0 ? 0 : depth[p] + 1;
        if (depth[j] > maxd) maxd = depth[j];
    }
    /* engine order: by depth (the kernel walks depth by depth), stable within a depth */
    std::vector<int> order; order.reserve(J);
    std::vector<int32_t> dstart(maxd + 2, 0);
    for (int d = 0; d <= maxd; ++d) { dstart[d] = (int32_t)order.size(); for (int j = 0; j < J; ++j) if (depth[j] == d) order.push_back(j); }
    dstart[maxd + 1] = J;
    for (int i = 0; i < J; ++i) e->k_perm[order[i]] = i;
    std::vector<int32_t> parent(J), type(J), dof(J);
    std::vector<double> axis(3 * J), origin(12 * J), mass(J), com(3 * J), msub(J);
    for (int i = 0; i < J; ++i) {
        const int j = order[i];
        parent[i] = t->parent[j] < 0 ? -1 : e->k_perm[t->parent[j]];
        e->k_hparent[i] = parent[i];
        type[i] = t->type[j]; dof[i] = t->type[j] == OSB_KIN_FIXED ? -1 : t->dof[j];
        double nrm = sqrt(t->axis[3 * j] * t->axis[3 * j] + t->axis[3 * j + 1] * t->axis[3 * j + 1] + t->axis[3 * j + 2] * t->axis[3 * j + 2]);
        if (!(nrm > 0)) nrm = 1.0;
        for (int k = 0; k < 3; ++k) { axis[3 * i + k] = t->axis[3 * j + k] / nrm; com[3 * i + k] = t->com[3 * j + k]; }
        for (int k = 0; k < 12; ++k) origin[12 * i + k] = t->origin[12 * j + k];
        mass[i] = t->mass[j];
    }
    /* subtree masses (constants of the tree), summed bottom-up in the order the kernel sums the first moments */
    for (int i = 0; i < J; ++i) msub[i] = mass[i];
    for (int i = J - 1; i >= 0; --i) if (parent[i] >= 0) msub[parent[i]] += msub[i];
    CUDA_OK(cudaSetDevice(e->device));
    if (e->k_dev) { cudaFree(e->k_dev); e->k_dev = nullptr; }
    const size_t ints = (size_t)(3 * J + maxd + 2), dbl = (size_t)(3 * J + 12 * J + J + 3 * J + J);
    const size_t ibytes = (ints * 4 + 15) & ~(size_t)15;
    CUDA_OK(cudaMalloc(&e->k_dev, ibytes + dbl * 8));
    e->k_parent = (int32_t*)e->k_dev; e->k_type = e->k_parent + J; e->k_dofmap = e->k_type + J; e->k_depth_start = e->k_dofmap + J;
    e->k_axis = (double*)(e->k_dev + ibytes); e->k_origin = e->k_axis + 3 * J; e->k_mass = e->k_origin + 12 * J; e->k_com = e->k_mass + J; e->k_msub = e->k_com + 3 * J;
    CUDA_OK(cudaMemcpy(e->k_parent, parent.data(), 4 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_type, type.data(), 4 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_dofmap, dof.data(), 4 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_depth_start, dstart.data(), 4 * (maxd + 2), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_axis, axis.data(), 8 * 3 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_origin, origin.data(), 8 * 12 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_mass, mass.data(), 8 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_com, com.data(), 8 * 3 * J, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(e->k_msub, msub.data(), 8 * J, cudaMemcpyHostToDevice));
    e->k_joints = J; e->k_dof = t->n_dof; e->k_depths = maxd + 1;
    return OSB_SUCCESS;
}

extern "C" int osb_produce_kinematics_device(osb_engine* e, const double* d_q, int64_t q_stride, const osb_kin_task_desc* tasks,
                                             int32_t n_tasks, double* d_blobs, int32_t batch, void* cuda_stream)
{
    if (!e || !d_q || !tasks || !d_blobs || batch < 1 || n_tasks < 1 || n_tasks > OSB_KIN_MAX_TASKS)
        return fail(OSB_E_INVALID, "bad arguments (1..8 tasks per call)%s", "");
    if (!e->k_dev) return fail(OSB_E_INVALID, "no kinematic tree: call osb_kinematics_create first%s", "");
    kin_params g;
    memset(&g, 0, sizeof(g));
    const int J = e->k_joints, n = e->d.n;
    g.n_joints = J; g.n_dof = n; g.n_depths = e->k_depths; g.n_tasks = n_tasks; g.batch = batch; g.stride = e->L.stride;
    g.parent = e->k_parent; g.type = e->k_type; g.dof = e->k_dofmap; g.depth_start = e->k_depth_start;
    g.axis = e->k_axis; g.origin = e->k_origin; g.mass = e->k_mass; g.com = e->k_com; g.msub = e->k_msub;
    g.q = d_q; g.q_stride = q_stride; g.blobs = d_blobs;
    g.smem_per_warp = (J * 15 + (J + 7) / 8 + 1) & ~1;   /* doubles: T [J][12], first moments [J][3], parents (one byte each) */
    auto path = [&](int joint, uint64_t* bits) { for (int j = joint; j >= 0; j = e->k_hparent[j]) bits[j >> 6] |= 1ull << (j & 63); };
    for (int i = 0; i < n_tasks; ++i) {
        const osb_kin_task_desc& T = tasks[i];
        kin_task& K = g.task[i];
        const int rows = T.kind == OSB_KIN_COM ? 3 : 6;
        if (T.kind != OSB_KIN_COM && T.kind != OSB_KIN_CARTESIAN) return fail(OSB_E_INVALID, "unknown kinematic task kind%s", "");
        if (T.level < 0 || T.level >= e->d.n_levels || e->d.task_kind[T.level] != OSB_TASK_DENSE || T.row0 < 0 ||
            T.row0 + rows > e->d.m[T.level])
            return fail(OSB_E_INVALID, "kinematic task: destination rows outside an OSB_TASK_DENSE level%s", "");
        K.kind = T.kind; K.lambda = T.lambda; K.orientation_gain = T.orientation_gain;
        K.off_A = e->L.A[T.level] + T.row0 * n; K.off_b = e->L.b[T.level] + T.row0;
        K.desired = T.desired; K.desired_stride = T.desired_stride; K.desired_vel = T.desired_vel; K.desired_vel_stride = T.desired_vel_stride;
        K.distal = K.base = -1;
        if (T.kind == OSB_KIN_CARTESIAN) {
            if (T.distal < 0 || T.distal >= J || T.base >= J) return fail(OSB_E_INVALID, "kinematic task: link index out of range%s", "");
            K.distal = e->k_perm[T.distal]; path(K.distal, K.anc_distal);
            if (T.base >= 0) { K.base = e->k_perm[T.base]; path(K.base, K.anc_base); }
        } else g.need_com = 1;
    }
    CUDA_OK(cudaSetDevice(e->device));
    const size_t smem = (size_t)g.smem_per_warp * KIN_WARPS * sizeof(double);
    if (smem > 48 * 1024) CUDA_OK(cudaFuncSetAttribute((const void*)sot_kinematics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sot_kinematics_kernel<<<(unsigned)((batch + KIN_WARPS - 1) / KIN_WARPS), 32 * KIN_WARPS, smem, (cudaStream_t)cuda_stream>>>(g);
    CUDA_OK(cudaGetLastError());
    e->launches += 1;
    return OSB_SUCCESS;
}

/* ---- state-input path --------------------------------------------------------------------------- */
extern "C" int osb_states_configure(osb_engine* e, const osb_states_config* c)
{
    if (!e || !c || !c->blob_template || c->n_tasks < 0 || c->n_tasks > OSB_KIN_MAX_TASKS)
        return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (e->es != 8) return fail(OSB_E_INVALID, "the state-input path is fp64 only%s", "");
    if (c->n_tasks > 0 && !e->k_dev) return fail(OSB_E_INVALID, "kinematic tasks need osb_kinematics_create first%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    const int n = e->d.n, stride = e->L.stride;
    /* one device block: template | q_ref q_min q_max q_neutral qdot_max | per task: desired (12), desired_vel (6) */
    const size_t total = (size_t)stride + 5 * (size_t)n + (size_t)c->n_tasks * 18;
    if (e->s_dev) { cudaFree(e->s_dev); e->s_dev = nullptr; }
    CUDA_OK(cudaMalloc(&e->s_dev, total * sizeof(double)));
    if (!e->s_q) CUDA_OK(cudaMalloc(&e->s_q, (size_t)e->max_batch * n * sizeof(double)));
    std::vector<double> host(total, 0.0);
    memcpy(host.data(), c->blob_template, sizeof(double) * stride);
    const double* vecs[5] = {c->q_ref, c->q_min, c->q_max, c->q_neutral, c->qdot_max};
    for (int v = 0; v < 5; ++v) if (vecs[v]) memcpy(host.data() + stride + (size_t)v * n, vecs[v], sizeof(double) * n);
    memset(&e->s_posture, 0, sizeof(e->s_posture));
    double* base = e->s_dev + stride;
    e->s_posture.q = e->s_q; e->s_posture.q_stride = n;
    e->s_posture.q_ref = c->q_ref ? base : nullptr; e->s_posture.q_ref_stride = 0;
    e->s_posture.q_min = c->q_min ? base + n : nullptr; e->s_posture.q_max = c->q_max ? base + 2 * n : nullptr;
    e->s_posture.q_neutral = c->q_neutral ? base + 3 * n : nullptr; e->s_posture.qdot_max = c->qdot_max ? base + 4 * n : nullptr;
    e->s_posture.lambda = c->lambda; e->s_posture.bound_scaling = c->bound_scaling; e->s_posture.dT = c->dT;
    e->s_posture.postural_level = c->postural_level; e->s_posture.joint_limit_set = c->joint_limit_set;
    e->s_posture.velocity_limit_set = c->velocity_limit_set;
    for (int t = 0; t < c->n_tasks; ++t) {
        e->s_tasks[t] = c->tasks[t];
        double* slot = base + 5 * n + (size_t)t * 18;
        const int nd = c->tasks[t].kind == OSB_KIN_COM ? 3 : 12, nv = c->tasks[t].kind == OSB_KIN_COM ? 3 : 6;
        if (c->tasks[t].desired) { memcpy(host.data() + (slot - e->s_dev), c->tasks[t].desired, sizeof(double) * nd); e->s_tasks[t].desired = slot; }
        if (c->tasks[t].desired_vel) { memcpy(host.data() + (slot + 12 - e->s_dev), c->tasks[t].desired_vel, sizeof(double) * nv); e->s_tasks[t].desired_vel = slot + 12; }
        e->s_tasks[t].desired_stride = 0; e->s_tasks[t].desired_vel_stride = 0;
    }
    e->s_ntasks = c->n_tasks;
    CUDA_OK(cudaMemcpy(e->s_dev, host.data(), total * sizeof(double), cudaMemcpyHostToDevice));
    /* every instance blob starts as the template; the producers overwrite the q-dependent fields each cycle */
    const long long cells = (long long)e->max_batch * stride;
    sot_template_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, e->stream>>>((double*)e->d_blobs, e->s_dev, stride, cells);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(e->stream));
    e->s_ready = 1;
    e->ps[0].tag = OSB_TAG_STATES; e->ps[1].tag = OSB_TAG_NONE;
    return OSB_SUCCESS;
}

extern "C" int osb_produce_posture_device(osb_engine* e, const osb_posture_inputs* in, double* d_blobs, int32_t batch, void* cuda_stream);

static int states_enqueue(osb_engine* e, osb_pipe_slot& ps, const double* q, int32_t batch, double* x, int32_t* status, bool fence)
{
    const size_t n = (size_t)e->d.n, stride = (size_t)e->stride_e * e->es;
    /* chunks over the three engines of the GPU like osb_solve_host: joint positions of chunk c+1 cross PCIe while the
     * producers and the solver work on chunk c (two compute streams) and the solutions of chunk c-1 travel back */
    /* (measured at 4096 states: the blocking call 0.348 ms in one chunk, 0.331 in two, 0.363 in three -- the second chunk's
     * producers run under the first chunk's solve; with two calls in flight the overlap comes from the other call and one
     * chunk is best: 0.200 against 0.212 ms) */
    int chunks = batch >= 32768 ? 4 : (batch >= (fence ? 2048 : 16384) ? 2 : 1);
    if (const char* k = getenv("OSB_STATES_CHUNKS")) { chunks = atoi(k); if (chunks < 1) chunks = 1; if (chunks > OSB_PIPE_CHUNKS) chunks = OSB_PIPE_CHUNKS; }   /* development aid */
    {   /* the constant part of this slot's blobs (again after a blob-path or wire-format call used the slot) */
        int rc = ensure_template(e, ps, OSB_TAG_STATES, e->s_h2d);
        if (rc != OSB_SUCCESS) return rc;
    }
    const bool posture = e->s_posture.postural_level >= 0 || e->s_posture.joint_limit_set >= 0 || e->s_posture.velocity_limit_set >= 0;
    if (fence) {
        CUDA_OK(cudaEventRecord(e->ev_begin, e->stream));
        CUDA_OK(cudaStreamWaitEvent(e->s_h2d, e->ev_begin, 0));
        CUDA_OK(cudaStreamWaitEvent(e->s_k2, e->ev_begin, 0));
        CUDA_OK(cudaStreamWaitEvent(e->s_d2h, e->ev_begin, 0));
    }
    for (int c = 0; c < chunks; ++c) {
        const size_t lo = (size_t)batch * c / chunks, hi = (size_t)batch * (c + 1) / chunks, cnt = hi - lo;
        cudaStream_t sk = ((c + (&ps == &e->ps[1] ? 1 : 0)) & 1) ? e->s_k2 : e->stream;   /* (see host_enqueue) */
        CUDA_OK(cudaMemcpyAsync(ps.s_q + lo * n, q + lo * n, cnt * n * sizeof(double), cudaMemcpyHostToDevice, e->s_h2d));
        CUDA_OK(cudaEventRecord(ps.ev_h2d[c], e->s_h2d));
        CUDA_OK(cudaStreamWaitEvent(sk, ps.ev_h2d[c], 0));
        double* blobs = (double*)(ps.d_blobs + lo * stride);
        int rc = OSB_SUCCESS;
        if (posture) {
            osb_posture_inputs in = e->s_posture;
            in.q = ps.s_q + lo * n;
            rc = osb_produce_posture_device(e, &in, blobs, (int32_t)cnt, sk);
        }
        if (rc == OSB_SUCCESS && e->s_ntasks > 0)
            rc = osb_produce_kinematics_device(e, ps.s_q + lo * n, (int64_t)n, e->s_tasks, e->s_ntasks, blobs, (int32_t)cnt, sk);
        if (rc == OSB_SUCCESS)
            rc = launch_solve(e, blobs, (int32_t)cnt, ps.d_x + lo * n * e->es, ps.d_status + lo, nullptr, ps.d_iters + lo, sk);
        if (rc != OSB_SUCCESS) return rc;
        CUDA_OK(cudaEventRecord(ps.ev_k[c], sk));
        CUDA_OK(cudaStreamWaitEvent(e->s_d2h, ps.ev_k[c], 0));
        CUDA_OK(cudaMemcpyAsync(x + lo * n, ps.d_x + lo * n * e->es, cnt * n * sizeof(double), cudaMemcpyDeviceToHost, e->s_d2h));
        CUDA_OK(cudaMemcpyAsync(status + lo, ps.d_status + lo, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, e->s_d2h));
    }
    CUDA_OK(cudaEventRecord(ps.ev_end, e->s_d2h));
    return OSB_SUCCESS;
}

extern "C" int osb_solve_states_host(osb_engine* e, const double* q, int32_t batch, double* x, int32_t* status)
{
    if (!e || !q || !x || !status || batch < 1 || batch > e->max_batch) return fail(OSB_E_INVALID, "bad arguments (batch must be 1..max_batch)%s", "");
    if (!e->s_ready) return fail(OSB_E_INVALID, "call osb_states_configure first%s", "");
    if (e->ps_count > 0) return fail(OSB_E_INVALID, "submitted calls are still in flight: osb_solve_wait first%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    osb_pipe_slot* ps = nullptr;
    int rc = pipe_slot(e, 0, &ps);
    if (rc != OSB_SUCCESS) return rc;
    rc = states_enqueue(e, *ps, q, batch, x, status, true);
    if (rc != OSB_SUCCESS) return rc;
    CUDA_OK(cudaStreamWaitEvent(e->stream, e->ev_end, 0));
    CUDA_OK(cudaStreamSynchronize(e->stream));
    return OSB_SUCCESS;
}

/* submitted twin of osb_solve_states_host (osb_solve_wait completes it): the templated blobs are per slot */
extern "C" int osb_solve_states_submit(osb_engine* e, const double* q, int32_t batch, double* x, int32_t* status)
{
    if (!e || !q || !x || !status || batch < 1 || batch > e->max_batch) return fail(OSB_E_INVALID, "bad arguments (batch must be 1..max_batch)%s", "");
    if (!e->s_ready) return fail(OSB_E_INVALID, "call osb_states_configure first%s", "");
    osb_pipe_slot* ps = nullptr;
    int rc = submit_slot(e, &ps);
    if (rc != OSB_SUCCESS) return rc;
    rc = states_enqueue(e, *ps, q, batch, x, status, false);
    if (rc != OSB_SUCCESS) return rc;
    e->ps_count += 1;
    return OSB_SUCCESS;
}

extern "C" int osb_states_download_blobs(osb_engine* e, double* blobs, int32_t batch)
{
    if (!e || !blobs || batch < 1 || batch > e->max_batch) return fail(OSB_E_INVALID, "bad arguments%s", "");
    CUDA_OK(cudaSetDevice(e->device));
    CUDA_OK(cudaMemcpy(blobs, e->d_blobs, (size_t)batch * e->L.stride * sizeof(double), cudaMemcpyDeviceToHost));
    return OSB_SUCCESS;
}

extern "C" int osb_engine_get_info(osb_engine* e, osb_engine_info* out)
{
    if (!e || !out) return fail(OSB_E_INVALID, "bad arguments%s", "");
    out->device = e->device; out->sm_count = e->sm_count; out->warps_per_cta = e->block / 32;
    out->ctas_per_sm = e->ctas_per_sm; out->grid = e->max_batch;
    out->smem_per_cta = e->smem_bytes; out->blob_stride_bytes = (int64_t)e->stride_e * e->es;
    out->kernel_regs = e->kernel_regs; out->launches = e->launches;
    return OSB_SUCCESS;
}

extern "C" int osb_engine_device_buffers(osb_engine* e, double** d_blobs, double** d_x, int32_t** d_status)
{
    if (!e) return fail(OSB_E_INVALID, "bad arguments%s", "");
    if (d_blobs) *d_blobs = (double*)e->d_blobs;
    if (d_x) *d_x = (double*)e->d_x;
    if (d_status) *d_status = e->d_status;
    return OSB_SUCCESS;
}

extern "C" void* osb_engine_stream(osb_engine* e) { return e ? (void*)e->stream : nullptr; }

extern "C" const char* osb_engine_kernel_name(osb_engine* e)
{
    if (!e) return "";
    if (e->mw_kernel) {
        snprintf(e->kernel_name, sizeof(e->kernel_name), "mw_solve_kernel<%s,%dx%d warps>", e->es == 8 ? "f64" : "f32",
                 2 * e->mw.np, e->mw.nw);
        return e->kernel_name;
    }
    if (!e->rr_kernel) return e->es == 8 ? "sot_solve_kernel" : "sot_solve_kernel_f32";
    snprintf(e->kernel_name, sizeof(e->kernel_name), "rr_solve_kernel<%s,%dx%d>", e->es == 8 ? "f64" : "f32",
             2 * e->rr.np, e->rr.rpl);
    return e->kernel_name;
}

extern "C" double osb_measure_fp64_tflops(int32_t device, int32_t iters)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    sot_dfma_probe_kernel<<<blocks, threads>>>(out, 64); /* warm-up */
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        sot_dfma_probe_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}

extern "C" double osb_measure_fp32_tflops(int32_t device, int32_t iters)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    float* out = nullptr;
    if (cudaMalloc(&out, sizeof(float) * blocks * threads) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    sot_ffma_probe_kernel<<<blocks, threads>>>(out, 64); /* warm-up */
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        sot_ffma_probe_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -1.0;
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    return flops / (best * 1e-3) / 1e12;
}

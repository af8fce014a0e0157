/*
 * emu_abi.cpp -- TEST INFRASTRUCTURE ONLY: the subset of the C ABI (include/opensot_b200.h) that the C++
 * host layer (opensot_b200/host/) calls, served by the CPU warp emulator of the kernels (warp_emu.h)
 * instead of a GPU. Built as tests/emu/_bin/abi/libopensot_b200.so and put in front of the real library
 * with LD_LIBRARY_PATH by tests/test_host_cpp_emu.py, so that the host logic (piling, constructors,
 * options, level views) is checked against the reference on every CPU run. Nothing under opensot_b200/
 * links this; the GPU tests (tests/test_gpu_host_cpp.py) run the same binaries over the real library.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "opensot_b200.h"
#include "sot_params.h"

extern "C" int emu_solve_ex(const osb_stack_desc* d, const void* blobs, int32_t batch, void* x,
                            int32_t* status, void* x_levels, int32_t* iters, const osb_options* opt,
                            uint32_t active_mask, int variant, int order, int64_t* info, const double* eps_level);

struct osb_engine {
    osb_stack_desc d;
    osb_layout L;
    osb_options opt;
    uint32_t active_mask;
    double eps_level[OSB_MAX_LEVELS];
    int32_t max_batch;
    int rr;
};
static char g_err[256] = "";

extern "C" const char* osb_last_error(void) { return g_err; }
extern "C" const char* osb_version(void) { return "opensot_b200 C ABI over the CPU warp emulator (test infrastructure)"; }
extern "C" int osb_engine_create(osb_engine** out, const osb_stack_desc* desc, int32_t max_batch, int32_t)
{
    osb_layout L;
    if (!out || !desc || osb_blob_layout(desc, &L) != OSB_SUCCESS) { snprintf(g_err, sizeof(g_err), "invalid stack descriptor"); return OSB_E_INVALID; }
    osb_engine* e = (osb_engine*)calloc(1, sizeof(osb_engine));
    e->d = *desc; e->L = L; e->max_batch = max_batch;
    e->opt.tol_feas = 1e-12; e->opt.tol_rank = 1e-11;
    e->active_mask = (1u << desc->n_levels) - 1u;
    for (int k = 0; k < OSB_MAX_LEVELS; ++k) e->eps_level[k] = desc->eps;
    rr_workspace w;
    e->rr = rr_plan_workspace(desc, &L, &w, 8) >= 0;
    *out = e;
    return OSB_SUCCESS;
}
extern "C" void osb_engine_destroy(osb_engine* e) { free(e); }
extern "C" int osb_set_options(osb_engine* e, const osb_options* o)
{
    e->opt = *o;
    if (!(e->opt.tol_feas > 0)) e->opt.tol_feas = 1e-12;
    if (!(e->opt.tol_rank > 0)) e->opt.tol_rank = 1e-11;
    return OSB_SUCCESS;
}
extern "C" int osb_get_options(osb_engine* e, osb_options* o) { *o = e->opt; return OSB_SUCCESS; }
extern "C" int osb_set_active_levels(osb_engine* e, const int32_t* a)
{
    uint32_t m = 0;
    for (int k = 0; k < e->d.n_levels; ++k) if (a[k]) m |= 1u << k;
    if (!m) return OSB_E_INVALID;
    e->active_mask = m;
    return OSB_SUCCESS;
}
extern "C" int osb_set_eps_level(osb_engine* e, int32_t level, double eps)
{
    if (!e || level >= e->d.n_levels || !(eps >= 0.0)) return OSB_E_INVALID;
    for (int k = 0; k < e->d.n_levels; ++k) if (level < 0 || k == level) e->eps_level[k] = eps;
    if (level < 0) e->d.eps = eps;
    return OSB_SUCCESS;
}
extern "C" double osb_get_eps_level(osb_engine* e, int32_t level) { return (level >= 0 && level < e->d.n_levels) ? e->eps_level[level] : -1.0; }
extern "C" void* osb_host_alloc(int64_t bytes) { return bytes > 0 ? malloc((size_t)bytes) : nullptr; }
extern "C" void osb_host_free(void* p) { free(p); }
extern "C" const char* osb_engine_kernel_name(osb_engine* e) { return e->rr ? "emulated rr_solve_kernel" : "emulated sot_solve_kernel"; }
extern "C" int osb_solve_host(osb_engine* e, const double* blobs, int32_t batch, double* x, int32_t* status, double* x_levels)
{
    if (!e || batch < 1 || batch > e->max_batch) return OSB_E_INVALID;
    if (x_levels && e->opt.store_levels) memset(x_levels, 0, sizeof(double) * (size_t)batch * e->d.n_levels * e->d.n);
    int64_t info[4];
    const int rc = emu_solve_ex(&e->d, blobs, batch, x, status, x_levels, nullptr, &e->opt, e->active_mask, e->rr ? 2 : 0, 0, info, e->eps_level);
    return rc == 0 ? OSB_SUCCESS : OSB_E_INVALID;
}

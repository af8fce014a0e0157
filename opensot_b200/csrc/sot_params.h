/*
 * sot_params.h -- kernel parameter block and per-warp shared-memory workspace plan.
 * Shared between the host side of the C ABI (osb_capi.cu) and the kernels (sot_kernel.cu).
 */
#ifndef OSB_SOT_PARAMS_H_
#define OSB_SOT_PARAMS_H_

#include <stdint.h>
#include "opensot_b200.h"

/* Per-instance (= per warp) workspace, offsets in doubles from the warp's base. */
struct sot_workspace {
    int32_t NRmax;      /* rows of [x ; Cin x ; Cloc x]  (n + c_in + max c_loc)          */
    int32_t TRmax;      /* rows of the working row-matrix T (NRmax + task/aux rows)     */
    int32_t AUX;        /* length of the per-task-row vectors (res, rn2, piv)           */
    int32_t LDZ;        /* leading dim (odd) of ZB : (n + c_in) rows, col-major         */
    int32_t LDT;        /* leading dim (odd) of T  : TRmax rows, col-major              */
    int32_t LDR;        /* leading dim (odd) of R / reduced Hessian : n rows            */
    int32_t off_blob, off_ZB, off_T, off_R;
    int32_t off_xprev;  /* y_prev : NRmax (x_prev, Cin x_prev, Cloc x_prev)             */
    int32_t off_y, off_ylo, off_yhi, off_dy;    /* NRmax each                          */
    int32_t off_d, off_r, off_u, off_rinv, off_gz;  /* n (+1) each                     */
    int32_t off_res;    /* task residuals, row norms: 2 * aux rows                      */
    int32_t off_act;    /* int32[n+1]  : active list (row | side<<16)                   */
    int32_t off_where;  /* int32[NRmax]: position in the active list or -1             */
    int32_t off_piv;    /* int32[aux rows]: pivot column of a task row or -1           */
    int32_t off_bar;    /* mbarrier (8 bytes)                                           */
    int32_t total;      /* doubles per warp (even)                                      */
};

struct sot_params {
    osb_stack_desc d;
    osb_layout     L;
    sot_workspace  w;
    int32_t  max_iter;
    int32_t  store_levels;
    double   tol_feas;
    double   tol_rank;
    uint32_t active_mask;
    int32_t  batch;
    const double* blobs;
    double*  x;
    int32_t* status;
    double*  x_levels;
    int32_t* iters;
};

struct gather_params {
    osb_gather_src src[16];
    int32_t n_srcs;
    int32_t stride;
    int32_t batch;
    double* blobs;
};

#ifdef __cplusplus
extern "C" {
#endif
/* fills w from d/L; returns bytes of shared memory per warp */
int64_t sot_plan_workspace(const osb_stack_desc* d, const osb_layout* L, sot_workspace* w);
#ifdef __cplusplus
}
#endif

#endif

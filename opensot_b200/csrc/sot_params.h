/*
 * sot_params.h -- kernel parameter block and per-warp shared-memory workspace plan.
 * Shared between the host side of the C ABI (osb_capi.cu) and the kernels (sot_kernel.cuh).
 */
#ifndef OSB_SOT_PARAMS_H_
#define OSB_SOT_PARAMS_H_

#include <stdint.h>
#include "opensot_b200.h"

/* how a level's reduced Hessian is factored (structural: same for the whole batch) */
enum sot_level_path { SOT_PATH_TRIVIAL = 0, SOT_PATH_LQ = 1, SOT_PATH_DENSE = 2 };

#if defined(__CUDACC__)
#define SOT_HD __host__ __device__
#else
#define SOT_HD
#endif
static inline SOT_HD int sot_level_path_of(const osb_stack_desc* d, int k)
{
    if (d->m_reg == 0 && d->task_kind[k] == OSB_TASK_IDENTITY && d->weight_kind[k] == OSB_W_IDENTITY)
        return SOT_PATH_TRIVIAL;
    if (d->m_reg == 0 && d->task_kind[k] == OSB_TASK_DENSE && d->weight_kind[k] != OSB_W_DENSE)
        return SOT_PATH_LQ;
    return SOT_PATH_DENSE;
}

/* The register-row and CTA-per-stack kernels take one more case on the LQ path: a dense task with a dense SPD weight
 * W = C C' is the identity-weight task (C'A, C'b) -- same H = A'WA, same g = -A'Wb, same null space -- and is whitened in
 * place in the staged blob (rr_whiten, sot_kernel_rr.cuh). The generic kernel keeps forming the reduced Hessian. */
static inline SOT_HD int rr_level_path_of(const osb_stack_desc* d, int k)
{
    if (d->m_reg == 0 && d->task_kind[k] == OSB_TASK_DENSE && d->weight_kind[k] == OSB_W_DENSE) return SOT_PATH_LQ;
    return sot_level_path_of(d, k);
}

/*
 * Per-instance (= per warp) workspace, offsets in ELEMENTS (double or float, the engine's
 * precision) from the warp's base.
 * All matrices are ROW-major with one odd row stride LD >= n: a lane walking along its own
 * row touches consecutive words (immediate-offset LDS in unrolled loops), lanes reading one
 * column hit 32 different banks (odd stride), lanes reading one row are consecutive.
 */
struct sot_workspace {
    int32_t NRmax;      /* rows of y = [x ; Cin x ; Cloc x]  (n + c_in + max c_loc)        */
    int32_t LD;         /* odd row stride of ZB, T, R, X                                  */
    int32_t RR;         /* rows of R: G/Hz/Cholesky, working-set factor, LQ task rows     */
    int32_t XR;         /* rows of X: task/aux rows of the DENSE path (0 if unused)       */
    int32_t AUX;        /* length of the per-task-row vectors (res, rn2, piv)             */
    int32_t off_blob, off_ZB, off_T, off_R, off_X;
    int32_t off_yprev;  /* NRmax (x_prev, Cin x_prev, Cloc x_prev)                        */
    int32_t off_y, off_ylo, off_yhi, off_dy;    /* NRmax each                            */
    int32_t off_d, off_r, off_u, off_rinv, off_gz;  /* n + 2 each                        */
    int32_t off_dm;     /* n + 2 : candidate row masked to the free columns               */
    int32_t off_rowsq;  /* NRmax : squared row norms of N                                 */
    int32_t off_res;    /* 2 * AUX : task residuals, row norms                            */
    int32_t off_act;    /* int32[n+2]  : working set (row | side<<16)                     */
    int32_t off_where;  /* int32[NRmax]: position in the working set, -1 free, -2 parked  */
    int32_t off_piv;    /* int32[AUX]  : pivot column of a task row or -1                 */
    int32_t off_bar;    /* mbarrier (8 bytes)                                             */
    int32_t total;      /* doubles per warp (even)                                        */
};

/*
 * Workspace of the register-row kernel (sot_kernel_rr.cuh): the matrix being transformed lives
 * in the registers of its row-owning lanes; shared memory holds the staged blob, the basis
 * ZB = [Z ; Cin Z ; Cloc Z] (NRmax x LD, row-major, ABSOLUTE column indices: the live basis is
 * columns [z0, n)), one work region and the small vectors. The work region is, in turn, the
 * projected task rows M of a level (read into registers by the sweep, which then writes its
 * reflectors over them) next to the compact factor Lc, then (over M) the reduced Hessian G of
 * the level, then (over everything) the packed working-set factor R of the active-set loop.
 */
struct rr_workspace {
    int32_t np;         /* register pairs per row: the kernel instance (LD = 2 np >= n)       */
    int32_t rpl;        /* rows per lane: the kernel instance (32 rpl >= every row count)     */
    int32_t LD;
    int32_t NRmax;      /* n + c_in + max c_loc                                              */
    int32_t MR;         /* rows of M: max(c_eq, task rows of an LQ level)                    */
    int32_t LDL;        /* row stride of the compact factor Lc (>= widest leading block)     */
    int32_t AUX;        /* length of per-task-row vectors                                    */
    int32_t stage_from; /* first blob element staged in shared memory (the level blocks)     */
    int32_t off_blob, off_ZB, off_M, off_L, off_R;
    int32_t off_vec;    /* 8 x LD : d, dm, vm, p, cscale, u, r, rinv (res, beta alias d, dm)  */
    int32_t off_rowv;   /* 4 x 32 rpl : y_prev, lo, hi, |row|^2 of N                          */
    int32_t off_act, off_piv;                                    /* int32 [LD], [LD]         */
    int32_t off_bar;
    int32_t total;
};

/*
 * Workspace of the CTA-per-stack kernel (sot_kernel_mw.cuh): nw warps, thread t owns row t. Like
 * rr_workspace, plus common and per-warp private column vectors and a few scalars for the exchange
 * between the warps.
 */
struct mw_workspace {
    int32_t np, nw, LD, NRmax, MR, LDL, AUX, stage_from;
    int32_t off_blob, off_ZB, off_M, off_L;
    int32_t off_cvec;   /* 7 x LD common: p, cscale, u, rinv, raw/res, csn (2 LD: beta, gv before the loop) */
    int32_t off_pvec;   /* nw x 4 x LD private: d, dm, vm, oh                                              */
    int32_t off_rowv;   /* 4 x 32 nw : y_prev, lo, hi, |row|^2 of N                                        */
    int32_t off_act, off_piv, off_scal, off_bar;
    int32_t total;
};

struct sot_params {
    osb_stack_desc d;
    osb_layout     L;
    sot_workspace  w;
    rr_workspace   rr;
    mw_workspace   mw;
    int32_t  max_iter;
    int32_t  store_levels;
    double   tol_feas;
    double   tol_rank;
    double   eps_level[OSB_MAX_LEVELS];   /* absolute diagonal shift of level k (default d.eps; iHQP::setEpsRegularisation(eps, i)) */
    uint32_t active_mask;
    int32_t  batch;
    const void* blobs;      /* element type = the engine's precision (double or float) */
    void*    x;
    int32_t* status;
    void*    x_levels;
    int32_t* iters;
    /* fused all-gather (osb_solve_device_gather): every x row and status word is also stored into
     * the gather buffers of the peer GPUs, at global instance peer_first + inst */
    int32_t  n_peers;
    int64_t  peer_first;
    void*    peer_x[OSB_MAX_PEERS];
    int32_t* peer_status[OSB_MAX_PEERS];
};

struct gather_params {
    osb_gather_src src[16];
    int32_t n_srcs;
    int32_t stride;
    int32_t batch;
    double* blobs;
};

struct gather_barrier_params {
    uint32_t* flags[OSB_MAX_PEERS];   /* flag block of rank r (peer memory): word [src] = last epoch src reached */
    int32_t   rank, world;
    uint32_t  epoch;
    long long timeout_cycles;
    int32_t*  fail;
};

struct posture_params {
    osb_posture_inputs in;
    int32_t n, stride, batch;
    int32_t off_b, off_jl, off_ju, off_vl, off_vu;   /* blob offsets, -1 = producer skipped */
    double* blobs;
};

/* ---- Cartesian / CoM producers (sot_kinematics.cuh) ---- */
#define OSB_KIN_MAX_TASKS 8
#define OSB_KIN_MAX_JOINTS 128
struct kin_task {
    int32_t kind, distal, base, pad;        /* joints in the engine's depth-sorted order; base < 0: world */
    int32_t off_A, off_b;                   /* blob offsets of the first row of the task                  */
    double  lambda, orientation_gain;
    const double* desired;     int64_t desired_stride;
    const double* desired_vel; int64_t desired_vel_stride;
    uint64_t anc_distal[2], anc_base[2];    /* bit j: joint j lies on the path from the root to the link  */
};
struct kin_params {
    int32_t n_joints, n_dof, n_depths, n_tasks, batch, stride, need_com, smem_per_warp;
    const int32_t *parent, *type, *dof, *depth_start;
    const double *axis, *origin, *mass, *com, *msub;   /* msub: mass of every joint's subtree */
    const double* q; int64_t q_stride;
    double* blobs;
    kin_task task[OSB_KIN_MAX_TASKS];
};

#ifdef __cplusplus
extern "C" {
#endif
/* fills w from d/L; returns bytes of shared memory per warp */
int64_t sot_plan_workspace(const osb_stack_desc* d, const osb_layout* L, sot_workspace* w, int elem_size);
/* same for the register-row kernel; returns -1 when the shape is outside its instances (the
 * generic kernel then solves it) */
int64_t rr_plan_workspace(const osb_stack_desc* d, const osb_layout* L, rr_workspace* w, int elem_size);
/* same for the CTA-per-stack kernel (33..96 rows of y, or 39..64 variables); -1 outside its instances */
int64_t mw_plan_workspace(const osb_stack_desc* d, const osb_layout* L, mw_workspace* w, int elem_size);
#ifdef __cplusplus
}
#endif

#endif

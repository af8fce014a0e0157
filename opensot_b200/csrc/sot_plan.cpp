/*
 * sot_plan.cpp -- per-warp shared-memory workspace plans of the solver kernels (pure host code,
 * no CUDA): sot_plan_workspace for the generic kernel (sot_kernel.cuh), rr_plan_workspace for
 * the register-row kernel (sot_kernel_rr.cuh). Shared by the C ABI (osb_capi.cu) and by the
 * warp emulator of tests/emu.
 */
#include "sot_params.h"

extern "C" int64_t sot_plan_workspace(const osb_stack_desc* d, const osb_layout* L, sot_workspace* w, int elem_size)
{
    const int ipe = elem_size / 4;                 /* int32 per element */
    const int stride_e = (elem_size == 8) ? L->stride : ((L->stride + 3) & ~3);
    const int n = d->n;
    int cloc_max = 0, rr = n > d->c_eq ? n : d->c_eq, xr = 0, aux = n > d->c_eq ? n : d->c_eq;
    if (d->m_reg > aux) aux = d->m_reg;
    for (int k = 0; k < d->n_levels; ++k) {
        if (d->c_loc[k] > cloc_max) cloc_max = d->c_loc[k];
        const int m = (d->task_kind[k] == OSB_TASK_HESSIAN) ? 0 : d->m[k];
        if (m > aux) aux = m;
        const int path = sot_level_path_of(d, k);
        if (path == SOT_PATH_LQ) { /* G in rows [0, min(m,n)), task rows after it */
            const int need = (m < n ? m : n) + m;
            if (need > rr) rr = need;
        } else if (path == SOT_PATH_DENSE) {
            int x = 0;
            if (d->task_kind[k] == OSB_TASK_DENSE) x = m * (d->weight_kind[k] == OSB_W_DENSE ? 2 : 1);
            else if (d->task_kind[k] == OSB_TASK_IDENTITY) x = (d->weight_kind[k] == OSB_W_DENSE) ? n : 0;
            else x = n;
            x += d->m_reg;
            if (x > xr) xr = x;
        }
    }
    w->NRmax = n + d->c_in + cloc_max;
    w->LD = (n + (n & 1)) + ((6 - ((n + (n & 1)) & 3)) & 3); /* smallest LD >= n (+1 pad word if n is odd) with LD % 4 == 2: rows 16-byte aligned and
                                       * LD/2 odd, so 128-bit row walks of a quarter-warp do not
                                       * collide on shared-memory banks */
    w->RR = rr;
    w->XR = xr;
    w->AUX = aux;
    int cur = 0;
    auto take = [&](int count) { int at = cur; cur += count + (count & 1); return at; };
    w->off_blob = take(stride_e);
    /* Words of the blob that are dead once the warp has read them during start-up (Cin is
     * copied into ZB, lA/uA and the bound sets are folded into ylo/yhi): reused for vectors
     * that are first written after that point. */
    int pool = w->off_blob + (d->c_in > 0 ? L->Cin : (d->n_bound_sets > 0 ? L->l[0] : L->global_size));
    int pool_end = w->off_blob + (d->m_reg > 0 ? L->Ar : L->global_size);
    pool += (pool & 1);
    auto take_late = [&](int count) {
        count += (count & 1);
        if (pool + count <= pool_end) { int at = pool; pool += count; return at; }
        return take(count);
    };
    w->off_ZB = take((n + d->c_in) * w->LD);
    w->off_T = take(w->NRmax * w->LD);
    w->off_R = take(w->RR * w->LD);
    w->off_X = take(w->XR * w->LD);
    w->off_yprev = take(w->NRmax);
    w->off_y = take(w->NRmax);
    w->off_ylo = take(w->NRmax);
    w->off_yhi = take(w->NRmax);
    w->off_d = take_late(n + 2);
    w->off_dm = take_late(n + 2);
    w->off_rinv = take_late(n + 2);
    w->off_gz = take(n + 2);  /* written by the equality stage while deq is still being read   */
    w->off_r = w->off_gz;     /* r lives in the active-set loop, gz only before it            */
    w->off_rowsq = take(w->NRmax); /* (y_prev stays alive through the active-set loop: a full vertex recomputes the
                                    * iterate from it) */
    {   /* res/rn2 live outside the active-set loop, dy and u only inside it */
        const int nre = w->NRmax + (w->NRmax & 1), auxe = w->AUX + (w->AUX & 1);
        const int u_at = nre > auxe ? nre : auxe;
        const int size = (2 * auxe > u_at + n + 2) ? 2 * auxe : u_at + n + 2;
        w->off_res = take_late(size);
        w->off_dy = w->off_res;
        w->off_u = w->off_res + u_at;
    }
    w->off_act = take((n + 2 + ipe - 1) / ipe + 1);
    w->off_where = take((w->NRmax + ipe - 1) / ipe + 1);
    w->off_piv = take((w->AUX + ipe - 1) / ipe + 1);
    cur += (cur & (ipe == 1 ? 3 : 1)) ? ((ipe == 1 ? 4 : 2) - (cur & (ipe == 1 ? 3 : 1))) : 0; /* 16-byte boundary */
    w->off_bar = take(8 / elem_size * 2);
    w->total = cur;
    return (int64_t)cur * elem_size;
}


extern "C" int64_t rr_plan_workspace(const osb_stack_desc* d, const osb_layout* L, rr_workspace* w, int elem_size)
{
    const int ipe = elem_size / 4;
    const int stride_e = (elem_size == 8) ? L->stride : ((L->stride + 3) & ~3);
    const int n = d->n;
    const int NC = 16;                  /* widest leading block the kernel instances handle */
    if (d->m_reg > 0) return -1;
    /* kernel instances: 30 or 38 columns per register row. Narrow stacks would waste most of a
     * row (measured: n = 12 runs faster on the generic kernel), wide ones do not fit. */
    int np = (n >= 20 && n <= 30) ? 15 : ((n > 30 && n <= 38) ? 19 : 0);
    if (!np) return -1;
    int cloc_max = 0, mr = d->c_eq, aux = n > d->c_eq ? n : d->c_eq, any_lq = 0;
    int ldl = (d->c_eq < n ? d->c_eq : n) + 2;
    for (int k = 0; k < d->n_levels; ++k) {
        const int path = rr_level_path_of(d, k);
        if (path == SOT_PATH_DENSE) return -1;
        if (elem_size == 4 && d->weight_kind[k] == OSB_W_DENSE) return -1;   /* only the fp64 instances carry rr_whiten */
        if (d->c_loc[k] > cloc_max) cloc_max = d->c_loc[k];
        if (path == SOT_PATH_LQ) {
            const int rmax = d->m[k] < n ? d->m[k] : n;
            if (rmax + 1 > NC) return -1;
            any_lq = 1;
            if (d->m[k] > mr) mr = d->m[k];
            if (d->m[k] > aux) aux = d->m[k];
        }
    }
    if (any_lq && ldl < NC) ldl = NC;
    ldl += (ldl & 1);
    const int nrmax = n + d->c_in + cloc_max;
    if (nrmax > 64 || mr > 2 * np) return -1;     /* per-task-row vectors hold LD entries */
    int rpl = (nrmax <= 32 && mr <= 32 && np == 15) ? 1 : 2;
    w->np = np; w->rpl = rpl; w->LD = 2 * np; w->NRmax = nrmax;
    w->MR = mr; w->LDL = ldl;
    w->AUX = aux;
    const int LD = w->LD;
    int cur = 0;
    auto take = [&](int count) { int at = cur; cur += count + (count & 1); return at; };
    w->stage_from = L->level_begin[0] & (elem_size == 8 ? ~1 : ~3);   /* 16-byte aligned */
    w->off_blob = take(stride_e - w->stage_from);
    w->off_ZB = take(nrmax * LD);
    {
        int msz = mr * LD;
        if (any_lq && msz < (NC + 2) * (NC + 1)) msz = (NC + 2) * (NC + 1);   /* reduced Hessian + gradient row */
        msz += (msz & 1);
        const int rp = n * (n + 1) / 2 + 2, both = msz + mr * ldl;
        w->off_M = take(both > rp ? both : rp);
        w->off_L = w->off_M + msz;
        w->off_R = w->off_M;
    }
    w->off_vec = take(8 * LD);
    w->off_rowv = take(4 * 32 * rpl);
    w->off_act = take((LD + ipe - 1) / ipe + 1);
    w->off_piv = w->off_act;            /* pivots live in the sweep, the working set in the active-set loop */
    cur += (cur & (ipe == 1 ? 3 : 1)) ? ((ipe == 1 ? 4 : 2) - (cur & (ipe == 1 ? 3 : 1))) : 0;
    w->off_bar = take(8 / elem_size * 2);
    w->total = cur;
    return (int64_t)cur * elem_size;
}


extern "C" int64_t mw_plan_workspace(const osb_stack_desc* d, const osb_layout* L, mw_workspace* w, int elem_size)
{
    const int ipe = elem_size / 4;
    const int stride_e = (elem_size == 8) ? L->stride : ((L->stride + 3) & ~3);
    const int n = d->n;
    const int NC = 16;
    if (d->m_reg > 0) return -1;
    /* kernel instances: 38 or 64 columns per register row, 2 or 3 warps (64 or 96 rows of y) */
    const int np = (n >= 20 && n <= 38) ? 19 : ((n > 38 && n <= 64) ? 32 : 0);
    if (!np) return -1;
    int cloc_max = 0, mr = d->c_eq, aux = n > d->c_eq ? n : d->c_eq, any_lq = 0;
    int ldl = (d->c_eq < n ? d->c_eq : n) + 2;
    for (int k = 0; k < d->n_levels; ++k) {
        const int path = rr_level_path_of(d, k);
        if (path == SOT_PATH_DENSE) return -1;
        if (elem_size == 4 && d->weight_kind[k] == OSB_W_DENSE) return -1;   /* only the fp64 instances carry rr_whiten */
        if (d->c_loc[k] > cloc_max) cloc_max = d->c_loc[k];
        if (path == SOT_PATH_LQ) {
            const int rmax = d->m[k] < n ? d->m[k] : n;
            if (rmax + 1 > NC) return -1;
            any_lq = 1;
            if (d->m[k] > mr) mr = d->m[k];
            if (d->m[k] > aux) aux = d->m[k];
        }
    }
    if (any_lq && ldl < NC) ldl = NC;
    ldl += (ldl & 1);
    const int nrmax = n + d->c_in + cloc_max;
    const int LD = 2 * np;
    const int LDR = (np & 1) ? LD : LD + 2;     /* row stride of the basis and the task block: MW_LDR in sot_kernel_mw.cuh */
    if (mr > 32 || mr > LD) return -1;          /* phase 1 of a sweep runs on one warp; per-task-row vectors hold LD entries */
    int nw = (nrmax + 31) / 32;
    if (nw < 2) nw = 2;
    if (nw > 3 || (np == 19 && nw > 2)) return -1;
    w->np = np; w->nw = nw; w->LD = LD; w->NRmax = nrmax; w->MR = mr; w->LDL = ldl; w->AUX = aux;
    int cur = 0;
    auto take = [&](int count) { int at = cur; cur += count + (count & 1); return at; };
    w->stage_from = L->level_begin[0] & (elem_size == 8 ? ~1 : ~3);
    w->off_blob = take(stride_e - w->stage_from);
    w->off_ZB = take(nrmax * LDR);
    {
        int msz = mr * LDR;
        if (any_lq && msz < (NC + 2) * (NC + 1)) msz = (NC + 2) * (NC + 1);
        msz += (msz & 1);
        const int rp = n * (n + 1) / 2 + 2, both = msz + mr * ldl;
        w->off_M = take(both > rp ? both : rp);
        w->off_L = w->off_M + msz;
    }
    w->off_cvec = take(7 * LD);
    w->off_pvec = take(nw * 4 * LD);
    w->off_rowv = take(4 * 32 * nw);
    w->off_act = take((LD + ipe - 1) / ipe + 1);
    w->off_piv = take((LD + ipe - 1) / ipe + 1);
    w->off_scal = take((2 * nw + 2) * (8 / elem_size));     /* slots are 8 bytes wide in either precision */
    cur += (cur & (ipe == 1 ? 3 : 1)) ? ((ipe == 1 ? 4 : 2) - (cur & (ipe == 1 ? 3 : 1))) : 0;
    w->off_bar = take(8 / elem_size * 2);
    w->total = cur;
    return (int64_t)cur * elem_size;
}

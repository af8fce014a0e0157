/*
 * osb_layout.c -- blob layout of one stack instance (pure C, no CUDA).
 *
 * The layout is the GPU counterpart of the reference's run-time piling
 * (include/OpenSoT/utils/Piler.h:102-163, src/tasks/Aggregated.cpp:113-132,
 * src/constraints/Aggregated.cpp:73-257): row offsets are fixed once per stack shape, so
 * "piling" child tasks of a level is writing their row blocks at consecutive offsets.
 *
 * Order inside a blob:  [global block][level 0 block] ... [level L-1 block]
 *   global block : Ceq deq | Cin lA uA | l_0 u_0 ... | Ar br wr
 *   level block  : A (or H) | b (or g) | w (or W) | c | Cl lCl uCl
 * Every block starts on an even double (16 B) so it can be fetched with cp.async.bulk.
 */
#include "opensot_b200.h"
#include <string.h>

static int32_t put(int32_t* cursor, int32_t count)
{
    int32_t at = *cursor;
    *cursor += count;
    return at;
}
static void pad_even(int32_t* cursor) { *cursor += (*cursor & 1); }

int osb_blob_layout(const osb_stack_desc* d, osb_layout* L)
{
    if (!d || !L) return OSB_E_INVALID;
    if (d->n < 1 || d->n > OSB_MAX_N) return OSB_E_TOO_LARGE;
    if (d->n_levels < 1 || d->n_levels > OSB_MAX_LEVELS) return OSB_E_INVALID;
    if (d->c_eq < 0 || d->c_in < 0 || d->m_reg < 0) return OSB_E_INVALID;
    if (d->n_bound_sets < 0 || d->n_bound_sets > OSB_MAX_BOUND_SETS) return OSB_E_INVALID;
    if (!(d->eps >= 0.0)) return OSB_E_INVALID;
    if (d->reg_weight_kind != OSB_W_IDENTITY && d->reg_weight_kind != OSB_W_DIAG)
        return OSB_E_INVALID;

    memset(L, 0xff, sizeof(*L)); /* every offset = -1 */
    const int32_t n = d->n;
    int32_t cur = 0;

    L->global_begin = cur;
    if (d->c_eq > 0) { L->Ceq = put(&cur, d->c_eq * n); L->deq = put(&cur, d->c_eq); }
    if (d->c_in > 0) {
        L->Cin = put(&cur, d->c_in * n);
        L->lA = put(&cur, d->c_in);
        L->uA = put(&cur, d->c_in);
    }
    for (int s = 0; s < d->n_bound_sets; ++s) { L->l[s] = put(&cur, n); L->u[s] = put(&cur, n); }
    if (d->m_reg > 0) {
        L->Ar = put(&cur, d->m_reg * n);
        L->br = put(&cur, d->m_reg);
        if (d->reg_weight_kind == OSB_W_DIAG) L->wr = put(&cur, d->m_reg);
    }
    pad_even(&cur);
    L->global_size = cur - L->global_begin;

    for (int k = 0; k < d->n_levels; ++k) {
        const int32_t m = d->m[k];
        if (m < 0 || d->c_loc[k] < 0) return OSB_E_INVALID;
        if (n + d->c_in + d->c_loc[k] + m + d->m_reg > OSB_MAX_ROWS) return OSB_E_TOO_LARGE;
        L->level_begin[k] = cur;
        switch (d->task_kind[k]) {
        case OSB_TASK_DENSE:
            L->A[k] = put(&cur, m * n);
            L->b[k] = put(&cur, m);
            break;
        case OSB_TASK_IDENTITY:
            if (m != n) return OSB_E_INVALID;
            L->b[k] = put(&cur, n);
            break;
        case OSB_TASK_HESSIAN:
            if (k != d->n_levels - 1) return OSB_E_INVALID;
            if (d->weight_kind[k] != OSB_W_IDENTITY) return OSB_E_INVALID;
            L->A[k] = put(&cur, n * n);
            L->b[k] = put(&cur, n);
            break;
        default:
            return OSB_E_INVALID;
        }
        switch (d->weight_kind[k]) {
        case OSB_W_IDENTITY: break;
        case OSB_W_DIAG:  L->w[k] = put(&cur, m); break;
        case OSB_W_DENSE: L->w[k] = put(&cur, m * m); break;
        default: return OSB_E_INVALID;
        }
        if (d->has_c[k]) L->c[k] = put(&cur, n);
        if (d->c_loc[k] > 0) {
            L->Cl[k] = put(&cur, d->c_loc[k] * n);
            L->lCl[k] = put(&cur, d->c_loc[k]);
            L->uCl[k] = put(&cur, d->c_loc[k]);
        }
        pad_even(&cur);
        L->level_size[k] = cur - L->level_begin[k];
    }
    L->stride = cur;
    return OSB_SUCCESS;
}

void osb_struct_sizes(int32_t out[6])
{
    out[0] = (int32_t)sizeof(osb_stack_desc);
    out[1] = (int32_t)sizeof(osb_layout);
    out[2] = (int32_t)sizeof(osb_options);
    out[3] = (int32_t)sizeof(osb_engine_info);
    out[4] = (int32_t)sizeof(osb_gather_src);
    out[5] = (int32_t)sizeof(osb_posture_inputs);
}

/*
 * opensot_b200_synth.h -- deterministic synthetic stack generator (benchmark / test data).
 *
 * One generator shared bit-for-bit by bench.py and the tests (through libopensot_b200.so:
 * osb_synth_shape / osb_synth_fill) and by the oracle side (the reference driver and the C
 * restatement include this header), so that every arm sees identical inputs. Distributions follow SURVEY.md section 8(d): task Jacobians ~ N(0,1),
 * errors 0.01/0.02*N(0,1), level-1 weight blkdiag(0.1 I6, I6) (the "0.1*l_wrist + r_wrist"
 * of examples/cpp/coman_ik.cpp:442), velocity bounds +-0.02 intersected with a joint-limit
 * like box that always contains 0 (constraints/velocity/JointLimits.cpp:47-52).
 *
 * Shapes:  C29, C35f, C38c, C35h  + small shapes for unit tests.
 */
#ifndef OSB_SYNTH_H_
#define OSB_SYNTH_H_

#include <math.h>
#include <stdint.h>
#include <string.h>
#include "opensot_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct synth_rng { uint64_t s; } synth_rng;

static inline uint64_t synth_next(synth_rng* r)
{
    uint64_t z = (r->s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline void synth_seed(synth_rng* r, uint64_t seed, uint64_t instance)
{
    r->s = seed * 0xd1342543de82ef95ULL + instance * 0x2545f4914f6cdd1dULL + 0x1234567ULL;
    (void)synth_next(r); (void)synth_next(r);
}
static inline double synth_uniform(synth_rng* r) /* (0,1) */
{
    return ((double)(synth_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double synth_normal(synth_rng* r)
{
    double u1 = synth_uniform(r), u2 = synth_uniform(r);
    return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925286766559 * u2);
}

/* Fills desc for a named shape; returns 0, or -1 if the name is unknown. */
static inline int synth_shape(const char* name, double eps_factor, osb_stack_desc* d)
{
    memset(d, 0, sizeof(*d));
    d->eps = eps_factor * OSB_QPOASES_BASE_EPS;
    int n = 0, feet = 0, cin = 0, cloc = 0, m0 = 3;
    if (!strcmp(name, "C29")) { n = 29; }
    else if (!strcmp(name, "C35f")) { n = 35; feet = 12; }
    else if (!strcmp(name, "C38c")) { n = 38; cin = 26; }
    else if (!strcmp(name, "C35h")) { n = 35; cloc = 4; m0 = 6; }
    else if (!strcmp(name, "C12")) { n = 12; }            /* small, 3 levels (2/4/I)   */
    else if (!strcmp(name, "C7c")) { n = 7; cin = 3; }    /* 7-DoF arm + 3 one-sided   */
    else if (!strcmp(name, "C64c")) { n = 64; cin = 16; feet = 6; } /* maximum size    */
    else if (!strcmp(name, "C29w")) { n = 29; }           /* dense W on level 1        */
    else if (!strcmp(name, "C29r")) { n = 29; }           /* user regularisation task  */
    else if (!strcmp(name, "C29c")) { n = 29; }           /* linear cost terms c       */
    else if (!strcmp(name, "C20d")) { n = 20; cin = 4; }  /* weighted identity levels  */
    else if (!strcmp(name, "C16e")) { n = 16; feet = 4; cin = 5; cloc = 2; } /* everything */
    else return -1;
    d->n = n;
    d->n_levels = 3;
    d->m[0] = m0; d->m[1] = 12; d->m[2] = n;
    if (!strcmp(name, "C12")) { d->m[0] = 2; d->m[1] = 4; }
    if (!strcmp(name, "C7c")) { d->m[0] = 3; d->m[1] = 3; }
    d->task_kind[0] = OSB_TASK_DENSE; d->task_kind[1] = OSB_TASK_DENSE;
    d->task_kind[2] = OSB_TASK_IDENTITY;
    d->weight_kind[0] = OSB_W_IDENTITY; d->weight_kind[1] = OSB_W_DIAG;
    d->weight_kind[2] = OSB_W_IDENTITY;
    for (int k = 0; k < 3; ++k) d->c_loc[k] = cloc;
    if (!strcmp(name, "C29w")) d->weight_kind[1] = OSB_W_DENSE;
    if (!strcmp(name, "C29r")) { d->m_reg = 5; d->reg_weight_kind = OSB_W_DIAG; }
    if (!strcmp(name, "C29c")) { d->has_c[1] = 1; d->has_c[2] = 1; } /* c on a rank-deficient level is LP-like */
    if (!strcmp(name, "C20d")) { d->m[0] = 4; d->m[1] = 6; d->weight_kind[2] = OSB_W_DIAG; d->weight_kind[0] = OSB_W_DENSE; }
    if (!strcmp(name, "C16e")) { d->m[0] = 3; d->m[1] = 5; d->weight_kind[2] = OSB_W_DENSE; d->has_c[2] = 1; }
    d->c_eq = feet;
    d->c_in = cin;
    d->n_bound_sets = 2;
    return 0;
}

/*
 * Fills one instance blob. `tight` in [0,1] is the probability that one side of a joint's
 * limit box is tight (0.05 of the velocity bound), which controls how many bounds end up
 * active (about 3 at tight=0.05, about 12 at tight=0.35 for C29).
 */
static inline void synth_fill(const osb_stack_desc* d, const osb_layout* L, uint64_t seed,
                              uint64_t instance, double tight, double* blob)
{
    synth_rng r;
    synth_seed(&r, seed, instance);
    const int n = d->n;
    memset(blob, 0, sizeof(double) * (size_t)L->stride);
    for (int k = 0; k < d->n_levels; ++k) {
        const int m = d->m[k];
        const double bscale = (d->task_kind[k] == OSB_TASK_IDENTITY) ? 0.02 : 0.01;
        if (d->task_kind[k] == OSB_TASK_DENSE)
            for (int i = 0; i < m * n; ++i) blob[L->A[k] + i] = synth_normal(&r);
        for (int i = 0; i < m; ++i) blob[L->b[k] + i] = bscale * synth_normal(&r);
        if (d->weight_kind[k] == OSB_W_DIAG)
            for (int i = 0; i < m; ++i)
                blob[L->w[k] + i] = (d->task_kind[k] == OSB_TASK_IDENTITY) ? 0.5 + synth_uniform(&r)
                                                                           : ((i < m / 2) ? 0.1 : 1.0);
        if (d->weight_kind[k] == OSB_W_DENSE) { /* SPD: I + 0.3 * G G' / m */
            double* W = blob + L->w[k];
            double g[OSB_MAX_ROWS];
            for (int i = 0; i < m; ++i) g[i] = synth_normal(&r);
            for (int i = 0; i < m; ++i)
                for (int j = 0; j < m; ++j)
                    W[i * m + j] = (i == j ? 1.0 : 0.0) + 0.3 * g[i] * g[j] / (double)m;
        }
        if (d->has_c[k])
            for (int i = 0; i < n; ++i) blob[L->c[k] + i] = 0.001 * synth_normal(&r);
        if (d->c_loc[k] > 0) { /* one-sided hull-like rows: -inf <= a x <= |N(0,0.05)| */
            for (int i = 0; i < d->c_loc[k] * n; ++i) blob[L->Cl[k] + i] = synth_normal(&r);
            for (int i = 0; i < d->c_loc[k]; ++i) {
                blob[L->lCl[k] + i] = -OSB_INFTY;
                blob[L->uCl[k] + i] = fabs(0.05 * synth_normal(&r));
            }
        }
    }
    if (d->c_eq > 0) { /* contact rows, Ceq x = 0 */
        for (int i = 0; i < d->c_eq * n; ++i) blob[L->Ceq + i] = synth_normal(&r);
        for (int i = 0; i < d->c_eq; ++i) blob[L->deq + i] = 0.0;
    }
    if (d->c_in > 0) { /* collision / friction-cone like one-sided rows */
        for (int i = 0; i < d->c_in * n; ++i) blob[L->Cin + i] = synth_normal(&r);
        for (int i = 0; i < d->c_in; ++i) {
            blob[L->lA + i] = -OSB_INFTY;
            blob[L->uA + i] = fabs(0.05 * synth_normal(&r));
        }
    }
    if (d->n_bound_sets > 0) { /* set 0: velocity limits */
        for (int i = 0; i < n; ++i) { blob[L->l[0] + i] = -0.02; blob[L->u[0] + i] = 0.02; }
    }
    for (int s = 1; s < d->n_bound_sets; ++s) { /* joint-limit like box, 0 inside */
        for (int i = 0; i < n; ++i) {
            double lo = 0.05 * fabs(synth_normal(&r)) + 0.002;
            double hi = 0.05 * fabs(synth_normal(&r)) + 0.002;
            if (synth_uniform(&r) < tight) lo = 0.001 * synth_uniform(&r);
            if (synth_uniform(&r) < tight) hi = 0.001 * synth_uniform(&r);
            blob[L->l[s] + i] = -lo;
            blob[L->u[s] + i] = hi;
        }
    }
    if (d->m_reg > 0) {
        for (int i = 0; i < d->m_reg * n; ++i) blob[L->Ar + i] = 0.1 * synth_normal(&r);
        for (int i = 0; i < d->m_reg; ++i) blob[L->br + i] = 0.001 * synth_normal(&r);
        if (d->reg_weight_kind == OSB_W_DIAG)
            for (int i = 0; i < d->m_reg; ++i) blob[L->wr + i] = 0.5 + synth_uniform(&r);
    }
}

#ifdef __cplusplus
}
#endif
#endif /* OSB_SYNTH_H_ */

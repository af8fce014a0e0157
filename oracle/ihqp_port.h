/*
 * ihqp_port.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this
 * library; the product (opensot_b200/) never does.
 *
 * Parity status: PINNED. The restatement is checked (tests/test_oracle_port.py) against
 *   - the reference's own model-free known-answer tests (tests/golden/kat_backend.json,
 *     each entry citing the reference test file:line it restates), and
 *   - golden vectors produced by the UNMODIFIED reference compiled from /root/reference
 *     (oracle/_ref/ref_driver, iHQP + qpOASES through the BackEndFactory dlopen boundary),
 *     committed under tests/golden/ with the script that made them
 *     (tests/golden/make_golden.py).
 */
#ifndef IHQP_PORT_H_
#define IHQP_PORT_H_
#include <stdint.h>
#include "opensot_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* iHQP::computeCostFunction (src/solvers/iHQP.cpp:129-162): H = A'WA, g = -A'Wb + c.
 * A is m x n row-major; wkind/W as in osb_weight_kind; c may be NULL. H is n x n. */
void port_cost_function(int m, int n, const double* A, const double* b, int wkind,
                        const double* W, const double* c, double* H, double* g);

/* One strictly convex QP, the job of BackEnd::solve (src/solvers/QPOasesBackEnd.cpp:248-307):
 *   min 1/2 x'(H + eps I)x + g'x   s.t.  l <= x <= u,  lC <= C x <= uC
 * C is nc x n row-major, l/u may be NULL (no bounds). Returns an osb_status code. */
int port_qp_solve(int n, const double* H, const double* g, int nc, const double* C,
                  const double* lC, const double* uC, const double* l, const double* u,
                  double eps, double* x, int32_t* iters);

/* iHQP::solve (src/solvers/iHQP.cpp:263-358) for one instance blob.
 * active may be NULL (all levels active). x_levels may be NULL (else n_levels*n). */
int port_solve_stack(const osb_stack_desc* d, const double* blob, const int32_t* active,
                     double* x, double* x_levels, int32_t* iters);

/* batch loop over port_solve_stack; status[i] gets the osb_status of instance i */
int port_solve_batch(const osb_stack_desc* d, const double* blobs, int32_t batch,
                     const int32_t* active, double* x, double* x_levels, int32_t* status,
                     int32_t* iters);

/* synthetic inputs (oracle/synth.h) exposed for ctypes */
int  port_synth_shape(const char* name, double eps_factor, osb_stack_desc* d);
int  port_synth_fill(const osb_stack_desc* d, uint64_t seed, uint64_t first, int32_t count,
                     double tight, double* blobs);
int  port_blob_layout(const osb_stack_desc* d, osb_layout* L);

#ifdef __cplusplus
}
#endif
#endif

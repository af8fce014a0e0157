// ref_nhqp.cpp -- TEST INFRASTRUCTURE (never linked into the product): the reference's OWN nHQP front-end
// (src/solvers/nHQP.cpp, compiled unmodified by oracle/Makefile) driven over instance blobs. A separate
// translation unit because nHQP.h pulls in Eigen's BDCSVD, which the vendored Eigen 3.2.92 only compiles as
// C++17 (the rest of the driver is C++20).
//   nhqp_solve_blobs: per instance, Stack = the level tasks, bounds = one constraints::Aggregated of every
//   bound set and of the global rows; nHQP(stack, bounds, eps_factor, qpOASES); solve(x).
// out: per instance  ok, x[n]
#include <OpenSoT/solvers/nHQP.h>
#include <OpenSoT/constraints/Aggregated.h>

#include <cstdio>
#include <vector>

#include "ref_stack.h"
using namespace refstack;

extern "C" int nhqp_solve_blobs(const osb_stack_desc* d, const osb_layout* L, const double* blobs, int count,
                                double min_sv_ratio, int ab_regularisation, int ns_regularisation, double* out)
{
    const int n = d->n;
    for (int i = 0; i < count; ++i) {
        double* rec = out + (size_t)i * (1 + n);
        for (int j = 0; j <= n; ++j) rec[j] = 0.0;
        try {
            RefStack rs;
            build(rs, *d, *L, blobs + (size_t)i * L->stride, solvers::solver_back_ends::qpOASES, false);
            std::list<constraints::Aggregated::ConstraintPtr> all(rs.bounds.begin(), rs.bounds.end());
            if (rs.global_con) all.push_back(rs.global_con);
            auto bounds = std::make_shared<constraints::Aggregated>(all, n);
            bounds->update();
            solvers::iHQP::Stack stack = rs.stack->getStack();
            solvers::nHQP solver(stack, bounds, d->eps / OSB_QPOASES_BASE_EPS, solvers::solver_back_ends::qpOASES);
            solver.setMinSingularValueRatio(min_sv_ratio);
            solver.setPerformAbRegularization(ab_regularisation != 0);
            solver.setPerformSelectiveNullSpaceRegularization(ns_regularisation != 0);
            for (auto& t : stack) t->update();
            bounds->update();
            Eigen::VectorXd x(n); x.setZero();
            rec[0] = solver.solve(x) ? 1.0 : 0.0;
            for (int j = 0; j < n; ++j) rec[1 + j] = x[j];
        } catch (const std::exception& e) {
            fprintf(stderr, "nhqp instance %d: %s\n", i, e.what());
        }
    }
    return 0;
}

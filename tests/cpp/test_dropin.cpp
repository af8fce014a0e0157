// test_dropin.cpp -- TEST: the same C++ OpenSoT stack objects solved three ways.
//   (1) reference: one iHQP + qpOASES per stack                 (CPU, the oracle)
//   (2) ours     : OpenSoT::b200::BatchediHQP over all stacks   (one GPU solve through the C ABI)
//   (3) plugin   : unmodified iHQP driving libOpenSotBackEndODYS.so = B200BackEnd, level by level
// The stacks are built with the reference's own Math-of-Tasks operators (oracle/ref_stack.h).
// Prints one JSON line; the pytest wrapper (tests/test_gpu_host_cpp.py) asserts on it.
// usage: test_dropin <shape> <count> <tight> <seed> <n_plugin>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "BatchedHCOD.h"
#include "BatchediHQP.h"
#include "ref_stack.h"
#include "opensot_b200_synth.h"
#include "../../opensot_b200/csrc/osb_layout.c"

using namespace refstack;

int main(int argc, char** argv)
{
    if (argc != 6) { fprintf(stderr, "usage: test_dropin shape count tight seed n_plugin\n"); return 2; }
    std::cout.setstate(std::ios_base::failbit);
    osb_stack_desc d; osb_layout L;
    if (synth_shape(argv[1], 2e2, &d) || osb_blob_layout(&d, &L)) return 2;
    const int count = atoi(argv[2]); const double tight = atof(argv[3]);
    const uint64_t seed = strtoull(argv[4], 0, 10); const int n_plugin = atoi(argv[5]);
    const int n = d.n;
    std::vector<double> blob(L.stride);
    std::vector<RefStack> rs(count);
    std::vector<OpenSoT::AutoStack::Ptr> stacks;
    Eigen::MatrixXd Xref(n, count);
    int ref_fail = 0;
    for (int i = 0; i < count; ++i) {
        synth_fill(&d, &L, seed, (uint64_t)i, tight, blob.data());
        build(rs[i], d, L, blob.data(), solvers::solver_back_ends::qpOASES, true);
        Eigen::VectorXd x(n); x.setZero();
        rs[i].stack->update();
        if (!rs[i].solver->solve(x)) ++ref_fail;
        Xref.col(i) = x;
        stacks.push_back(rs[i].stack);
    }
    // (2) batched GPU front-end over the very same AutoStack objects
    double err_batched = -1; int bad_status = 0; bool ok_batched = false;
    try {
        OpenSoT::b200::BatchediHQP solver(stacks, 2e2, 0);
        Eigen::MatrixXd X;
        ok_batched = solver.updateAndSolve(X);
        for (int s : solver.getStatus()) bad_status += (s != 0);
        err_batched = (X - Xref).cwiseAbs().maxCoeff();
        // setActiveStack: level 1 off must equal the reference with the same level off
        if (d.n_levels == 3) {
            solver.setActiveStack(1, false);
            Eigen::MatrixXd X2; solver.updateAndSolve(X2);
            double e2 = 0;
            for (int i = 0; i < std::min(count, 8); ++i) {
                rs[i].solver->setActiveStack(1, false);
                Eigen::VectorXd x(n); rs[i].stack->update(); rs[i].solver->solve(x);
                rs[i].solver->activateAllStacks();
                e2 = std::max(e2, (X2.col(i) - x).cwiseAbs().maxCoeff());
            }
            printf("{\"active_level_err\": %.3e}\n", e2);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "BatchediHQP failed: %s\n", e.what());
    }
    // (2b) the HCOD-named surface over the same engine (damping -> eps)
    double err_hcod = -1;
    try {
        OpenSoT::b200::BatchedHCOD hcod(stacks, 0.0, 0);
        Eigen::MatrixXd X;
        for (auto& st : stacks) st->update();
        if (hcod.solve(X)) err_hcod = (X - Xref).cwiseAbs().maxCoeff();
    } catch (const std::exception& e) {
        fprintf(stderr, "BatchedHCOD failed: %s\n", e.what());
    }
    printf("{\"err_hcod\": %.3e}\n", err_hcod);
    // (3) unmodified iHQP + our plugin through BackEndFactory(ODYS)
    double err_plugin = 0; int plugin_fail = 0;
    for (int i = 0; i < std::min(n_plugin, count); ++i) {
        try {
            solvers::iHQP s(*rs[i].stack, 2e2, solvers::solver_back_ends::ODYS);
            Eigen::VectorXd x(n); x.setZero();
            rs[i].stack->update();
            if (!s.solve(x)) ++plugin_fail;
            err_plugin = std::max(err_plugin, (x - Xref.col(i)).cwiseAbs().maxCoeff());
        } catch (const std::exception& e) {
            fprintf(stderr, "plugin path failed: %s\n", e.what());
            ++plugin_fail;
        }
    }
    printf("{\"shape\": \"%s\", \"count\": %d, \"ref_fail\": %d, \"ok_batched\": %d, \"bad_status\": %d, "
           "\"err_batched\": %.3e, \"n_plugin\": %d, \"plugin_fail\": %d, \"err_plugin\": %.3e}\n",
           argv[1], count, ref_fail, ok_batched ? 1 : 0, bad_status, err_batched, std::min(n_plugin, count),
           plugin_fail, err_plugin);
    return 0;
}

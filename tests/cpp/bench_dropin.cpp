// bench_dropin.cpp -- HOST-PATH BENCH (test infrastructure: links the compiled reference as its host program).
// The "existing C++ stacks drop in unchanged" path, timed end to end on the SAME OpenSoT::AutoStack objects:
//   (1) reference loop      : for each stack  stack->update(); iHQP::solve(x)    (one host thread, hot-started)
//   (2) BatchediHQP         : updateAndSolve(X) = B x AutoStack::update() + piling into blobs (pack) +
//                             one osb_solve_host (H2D + kernel + D2H); update, pack and C-ABI time reported apart
//   (3) batch of one        : unmodified iHQP driving libOpenSotBackEndODYS.so (H2D + launch + D2H per level) and
//                             BatchediHQP(AutoStack&) against iHQP+qpOASES, latency per solve
// Prints one JSON line. usage: bench_dropin <shape> <count> <tight> <seed> <cycles>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "BatchediHQP.h"
#include "ref_stack.h"
#include "opensot_b200_synth.h"
#include "../../opensot_b200/csrc/osb_layout.c"

using namespace refstack;
typedef std::chrono::steady_clock Clock;
static double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char** argv)
{
    if (argc != 6 && argc != 7) { fprintf(stderr, "usage: bench_dropin shape count tight seed cycles [update_threads]\n"); return 2; }
    const int update_threads = argc == 7 ? atoi(argv[6]) : 1;    /* AutoStack::update() of the instances on this many host threads */
    std::cout.setstate(std::ios_base::failbit);
    osb_stack_desc d; osb_layout L;
    if (synth_shape(argv[1], 2e2, &d) || osb_blob_layout(&d, &L)) return 2;
    const int count = atoi(argv[2]); const double tight = atof(argv[3]);
    const uint64_t seed = strtoull(argv[4], 0, 10); const int cycles = atoi(argv[5]);
    const int n = d.n;
    std::vector<double> base((size_t)count * L.stride), blob(L.stride);
    std::vector<RefStack> rs(count);
    std::vector<OpenSoT::AutoStack::Ptr> stacks;
    for (int i = 0; i < count; ++i) {
        synth_fill(&d, &L, seed, (uint64_t)i, tight, base.data() + (size_t)i * L.stride);
        build(rs[i], d, L, base.data() + (size_t)i * L.stride, solvers::solver_back_ends::qpOASES, true);
        stacks.push_back(rs[i].stack);
    }
    synth_rng r; synth_seed(&r, seed ^ 0x77ULL, 0);
    auto drift = [&](int i) {        // time-coherent inputs, as a control loop produces them
        memcpy(blob.data(), base.data() + (size_t)i * L.stride, sizeof(double) * L.stride);
        for (int k = 0; k < d.n_levels; ++k)
            for (int j = 0; j < d.m[k]; ++j) blob[L.b[k] + j] *= 1.0 + 0.01 * synth_normal(&r);
        set_data(rs[i], d, L, blob.data());
    };
    // (1) the reference loop
    Eigen::MatrixXd Xref(n, count);
    double t_ref = 0; int ref_fail = 0;
    for (int c = -1; c < cycles; ++c) {
        for (int i = 0; i < count; ++i) drift(i);
        const auto t0 = Clock::now();
        for (int i = 0; i < count; ++i) {
            Eigen::VectorXd x(n);
            rs[i].stack->update();
            if (!rs[i].solver->solve(x)) ++ref_fail;
            Xref.col(i) = x;
        }
        if (c >= 0) t_ref += secs(t0, Clock::now());
    }
    // (2) the batched front-end on the same objects (same inputs as the last reference cycle first, for parity)
    OpenSoT::b200::BatchediHQP solver(stacks, 2e2, 0);
    solver.setStoreLevels(false);
    solver.setUpdateThreads(update_threads);
    Eigen::MatrixXd X;
    bool ok = solver.updateAndSolve(X);
    const double err = (X - Xref).cwiseAbs().maxCoeff();
    double t_all = 0, t_update = 0, t_pack = 0;
    for (int c = -2; c < cycles; ++c) {
        for (int i = 0; i < count; ++i) drift(i);
        const auto t0 = Clock::now();
        solver.update();
        const auto t1 = Clock::now();
        ok = solver.solve(X) && ok;
        const auto t2 = Clock::now();
        if (c >= 0) { t_all += secs(t0, t2); t_update += secs(t0, t1); t_pack += solver.getLastPackSeconds(); }
    }
    // (3) batch of one: latency per solve
    const int one_cycles = 200;
    double t_one_ref = 0, t_one_plugin = 0, t_one_batched = 0; int plugin_ok = 1;
    {
        Eigen::VectorXd x(n);
        for (int c = -5; c < one_cycles; ++c) {
            drift(0);
            const auto t0 = Clock::now();
            rs[0].stack->update(); rs[0].solver->solve(x);
            if (c >= 0) t_one_ref += secs(t0, Clock::now());
        }
        try {
            solvers::iHQP s(*rs[0].stack, 2e2, solvers::solver_back_ends::ODYS);
            for (int c = -5; c < one_cycles; ++c) {
                drift(0);
                const auto t0 = Clock::now();
                rs[0].stack->update(); if (!s.solve(x)) plugin_ok = 0;
                if (c >= 0) t_one_plugin += secs(t0, Clock::now());
            }
        } catch (const std::exception& e) { plugin_ok = 0; }
        OpenSoT::b200::BatchediHQP one(*rs[0].stack, 2e2);
        one.setStoreLevels(false);
        for (int c = -5; c < one_cycles; ++c) {
            drift(0);
            const auto t0 = Clock::now();
            rs[0].stack->update(); one.solve(x);
            if (c >= 0) t_one_batched += secs(t0, Clock::now());
        }
    }
    printf("\n{\"shape\": \"%s\", \"batch\": %d, \"cycles\": %d, \"kernel\": \"%s\", \"ok\": %d, \"ref_fail\": %d, \"err_vs_reference\": %.3e, "
           "\"update_threads\": %d, \"reference_loop_ms_per_cycle\": %.4f, \"reference_stacks_per_s_1thread\": %.1f, "
           "\"batched_ms_per_cycle\": %.4f, \"batched_update_ms\": %.4f, \"batched_pack_ms\": %.4f, \"batched_cabi_ms\": %.4f, "
           "\"batched_stacks_per_s\": %.1f, \"speedup_vs_reference_loop\": %.2f, "
           "\"one_reference_us\": %.2f, \"one_plugin_us\": %.2f, \"one_batched_us\": %.2f, \"plugin_ok\": %d}\n",
           argv[1], count, cycles, solver.getKernelName(), ok ? 1 : 0, ref_fail, err,
           update_threads, 1e3 * t_ref / cycles, count * cycles / t_ref,
           1e3 * t_all / cycles, 1e3 * t_update / cycles, 1e3 * t_pack / cycles, 1e3 * (t_all - t_update - t_pack) / cycles,
           count * cycles / t_all, t_ref / t_all,
           1e6 * t_one_ref / one_cycles, 1e6 * t_one_plugin / one_cycles, 1e6 * t_one_batched / one_cycles, plugin_ok);
    return 0;
}

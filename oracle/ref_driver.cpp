// ref_driver.cpp -- TEST / BENCH INFRASTRUCTURE (never linked into the product).
//
// Our own driver over the UNMODIFIED reference API (compiled from /root/reference by
// oracle/Makefile): builds an OpenSoT AutoStack out of GenericTask / GenericConstraint
// objects from a synthetic instance blob (oracle/synth.h), runs
//     AutoStack::update(); iHQP::solve()          (src/utils/AutoStack.cpp:385-393,
//                                                  src/solvers/iHQP.cpp:263-358)
// through the real BackEndFactory dlopen boundary (src/solvers/BackEndFactory.cpp:4-17)
// and dumps the solutions. Three uses:
//   solve  : golden vectors / live oracle for the parity tests
//   bench  : the reference's own CPU path timed on the host cores (bench.py cpu_baseline)
//   kat    : replays model-free known-answer tests of the reference through the build
//
// Usage:
//   ref_driver solve <shape> <eps_factor> <seed> <first> <count> <tight> <qpoases|eiquadprog> <out.bin>
//   ref_driver solveblob <desc.bin> <blobs.bin> <count> <qpoases|eiquadprog> <out.bin>
//   ref_driver bench <shape> <eps_factor> <seed> <count> <tight> <nproc> <cycles> [warmup_cycles]
//   ref_driver solveboth <desc.bin> <blobs.bin> <count> <out.bin>      cold AND hot-started qpOASES answers
//   ref_driver sequence <desc.bin> <blobs.bin> <count> <cycles> <seed> <out.bin> <blobs_out.bin>
//                                                      time-coherent cycles, persistent solver vs fresh solver
//   ref_driver kat
#include <OpenSoT/constraints/GenericConstraint.h>
#include <OpenSoT/constraints/TaskToConstraint.h>
#include <OpenSoT/solvers/iHQP.h>
#ifdef OSB_WITH_HCOD
#include <OpenSoT/solvers/HCOD.h>   // third oracle: soth's hierarchical complete orthogonal decomposition
#endif
#include <OpenSoT/tasks/GenericTask.h>
#include <OpenSoT/utils/AutoStack.h>

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "opensot_b200_synth.h"

using namespace OpenSoT;

// layout code is plain C shared with the product (a data format, not the algorithm)
#include "../opensot_b200/csrc/osb_layout.c"

#include "ref_stack.h"
using namespace refstack;

static bool g_hcod = false;   // "hcod": OpenSoT::solvers::HCOD (src/solvers/HCOD.cpp) instead of iHQP
static solvers::solver_back_ends parse_be(const char* s)
{
    g_hcod = !strcmp(s, "hcod");
    if (!strcmp(s, "eiquadprog")) return solvers::solver_back_ends::eiQuadProg;
    return solvers::solver_back_ends::qpOASES;
}

// out.bin: int32 n, L, count, pad ; then per instance: double ok, x[n], x_level[L][n]
static int solve_blobs(const osb_stack_desc& d, const osb_layout& L,
                       const std::vector<double>& blobs, int count,
                       solvers::solver_back_ends be, const char* out_path)
{
    FILE* f = fopen(out_path, "wb");
    if (!f) { perror("fopen"); return 2; }
    int32_t hdr[4] = {d.n, d.n_levels, count, 0};
    fwrite(hdr, sizeof(int32_t), 4, f);
    std::vector<double> rec(1 + d.n * (1 + d.n_levels));
    for (int i = 0; i < count; ++i) {
        const double* blob = blobs.data() + (size_t)i * L.stride;
        std::fill(rec.begin(), rec.end(), 0.0);
        bool ok = true;
        try {
            RefStack rs;
            build(rs, d, L, blob, be); // constructor solves every level once (cold start)
            Eigen::VectorXd x(d.n);
            x.setZero();
            rs.stack->update();
            if (g_hcod) {
#ifdef OSB_WITH_HCOD
                solvers::HCOD h(*rs.stack, 0.0);   // a true lexicographic solver: no eps, damping 0
                ok = h.solve(x);
                for (int j = 0; j < d.n; ++j) rec[1 + j] = x[j];
                rec[0] = ok ? 1.0 : 0.0;
                fwrite(rec.data(), sizeof(double), rec.size(), f);
                continue;
#else
                fprintf(stderr, "built without HCOD (make -C oracle hcod)\n"); fclose(f); return 2;
#endif
            }
            ok = rs.solver->solve(x); // hot-started from its own cold solution
            for (int j = 0; j < d.n; ++j) rec[1 + j] = x[j];
            for (int k = 0; k < d.n_levels; ++k) {
                solvers::BackEnd::Ptr bk;
                rs.solver->getBackEnd(k, bk);
                for (int j = 0; j < d.n; ++j) rec[1 + d.n * (1 + k) + j] = bk->getSolution()[j];
            }
        } catch (const std::exception& e) {
            ok = false;
        }
        rec[0] = ok ? 1.0 : 0.0;
        fwrite(rec.data(), sizeof(double), rec.size(), f);
    }
    fclose(f);
    return 0;
}

static int cmd_solve(int argc, char** argv)
{
    if (argc != 10) { fprintf(stderr, "solve: bad arguments\n"); return 2; }
    osb_stack_desc d; osb_layout L;
    if (synth_shape(argv[2], atof(argv[3]), &d) || osb_blob_layout(&d, &L)) {
        fprintf(stderr, "unknown shape %s\n", argv[2]); return 2; }
    const uint64_t seed = strtoull(argv[4], 0, 10);
    const int first = atoi(argv[5]), count = atoi(argv[6]);
    const double tight = atof(argv[7]);
    std::vector<double> blobs((size_t)count * L.stride);
    for (int i = 0; i < count; ++i)
        synth_fill(&d, &L, seed, (uint64_t)(first + i), tight, blobs.data() + (size_t)i * L.stride);
    return solve_blobs(d, L, blobs, count, parse_be(argv[8]), argv[9]);
}

static int cmd_solveblob(int argc, char** argv)
{
    if (argc != 7) { fprintf(stderr, "solveblob: bad arguments\n"); return 2; }
    osb_stack_desc d; osb_layout L;
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(&d, sizeof(d), 1, f) != 1) { fprintf(stderr, "cannot read desc\n"); return 2; }
    fclose(f);
    if (osb_blob_layout(&d, &L)) { fprintf(stderr, "bad desc\n"); return 2; }
    const int count = atoi(argv[4]);
    std::vector<double> blobs((size_t)count * L.stride);
    f = fopen(argv[3], "rb");
    if (!f || fread(blobs.data(), sizeof(double), blobs.size(), f) != blobs.size()) {
        fprintf(stderr, "cannot read blobs\n"); return 2; }
    fclose(f);
    return solve_blobs(d, L, blobs, count, parse_be(argv[5]), argv[6]);
}

// Level solutions of a solver object (BackEnd::getSolution of every level, iHQP.h:178-193)
static void final_x(const RefStack& rs, const osb_stack_desc& d, double* out)
{
    solvers::BackEnd::Ptr bk;
    rs.solver->getBackEnd(d.n_levels - 1, bk);
    for (int j = 0; j < d.n; ++j) out[j] = bk->getSolution()[j];
}

// out.bin: int32 n, L, count, 0 ; per instance: double ok_hot, ok_cold, x_hot[n], x_cold[n].
// "cold" = the answer of the constructor's initProblem cascade (iHQP.cpp:172-260: every level is
// set up and solved once, no prior working set); "hot" = a following solve() hot-started from it
// (iHQP.cpp:263-358), the production path. An instance is reference-stable when the two agree.
static int cmd_solveboth(int argc, char** argv)
{
    if (argc != 6) { fprintf(stderr, "solveboth: bad arguments\n"); return 2; }
    osb_stack_desc d; osb_layout L;
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(&d, sizeof(d), 1, f) != 1) { fprintf(stderr, "cannot read desc\n"); return 2; }
    fclose(f);
    if (osb_blob_layout(&d, &L)) { fprintf(stderr, "bad desc\n"); return 2; }
    const int count = atoi(argv[4]);
    std::vector<double> blobs((size_t)count * L.stride);
    f = fopen(argv[3], "rb");
    if (!f || fread(blobs.data(), sizeof(double), blobs.size(), f) != blobs.size()) {
        fprintf(stderr, "cannot read blobs\n"); return 2; }
    fclose(f);
    FILE* o = fopen(argv[5], "wb");
    if (!o) { perror("fopen"); return 2; }
    int32_t hdr[4] = {d.n, d.n_levels, count, 0};
    fwrite(hdr, sizeof(int32_t), 4, o);
    std::vector<double> rec(2 + 2 * d.n);
    for (int i = 0; i < count; ++i) {
        std::fill(rec.begin(), rec.end(), 0.0);
        try {
            RefStack rs;
            build(rs, d, L, blobs.data() + (size_t)i * L.stride, solvers::solver_back_ends::qpOASES);
            rec[1] = 1.0;
            final_x(rs, d, rec.data() + 2 + d.n);
            Eigen::VectorXd x(d.n); x.setZero();
            rs.stack->update();
            rec[0] = rs.solver->solve(x) ? 1.0 : 0.0;
            for (int j = 0; j < d.n; ++j) rec[2 + j] = x[j];
        } catch (const std::exception& e) { /* constructor threw: cold start failed (iHQP.cpp:24) */ }
        fwrite(rec.data(), sizeof(double), rec.size(), o);
    }
    fclose(o);
    return 0;
}

// Time-coherent sequence (A_k += 0.01 N(0,1) per cycle): `count` persistent solver
// objects are driven through `cycles` control cycles, hot-started from their own previous cycle as
// in production; beside each, a FRESH solver is built on the same inputs (cold start). Writes the
// blobs of every (cycle, instance) so that the GPU path is fed bit-identical inputs.
// out.bin: int32 n, cycles, count, 0 ; per (cycle, instance): ok_hot, ok_cold, x_hot[n], x_cold[n]
static int cmd_sequence(int argc, char** argv)
{
    if (argc != 9) { fprintf(stderr, "sequence: bad arguments\n"); return 2; }
    osb_stack_desc d; osb_layout L;
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(&d, sizeof(d), 1, f) != 1) { fprintf(stderr, "cannot read desc\n"); return 2; }
    fclose(f);
    if (osb_blob_layout(&d, &L)) return 2;
    const int count = atoi(argv[4]), cycles = atoi(argv[5]);
    const uint64_t seed = strtoull(argv[6], 0, 10);
    std::vector<double> blobs((size_t)count * L.stride);
    f = fopen(argv[3], "rb");
    if (!f || fread(blobs.data(), sizeof(double), blobs.size(), f) != blobs.size()) return 2;
    fclose(f);
    FILE* o = fopen(argv[7], "wb"); FILE* ob = fopen(argv[8], "wb");
    if (!o || !ob) { perror("fopen"); return 2; }
    int32_t hdr[4] = {d.n, cycles, count, 0};
    fwrite(hdr, sizeof(int32_t), 4, o);
    std::vector<RefStack> rs(count);
    std::vector<char> alive(count, 1);
    for (int i = 0; i < count; ++i) {
        try { build(rs[i], d, L, blobs.data() + (size_t)i * L.stride, solvers::solver_back_ends::qpOASES); }
        catch (const std::exception&) { alive[i] = 0; }
    }
    synth_rng r; synth_seed(&r, seed ^ 0x5e9ULL, 0);
    std::vector<double> rec(2 + 2 * d.n);
    Eigen::VectorXd x(d.n);
    for (int c = 0; c < cycles; ++c)
        for (int i = 0; i < count; ++i) {
            double* blob = blobs.data() + (size_t)i * L.stride;
            for (int k = 0; k < d.n_levels; ++k)
                if (d.task_kind[k] == OSB_TASK_DENSE)
                    for (int j = 0; j < d.m[k] * d.n; ++j) blob[L.A[k] + j] += 0.01 * synth_normal(&r);
            fwrite(blob, sizeof(double), L.stride, ob);
            std::fill(rec.begin(), rec.end(), 0.0);
            if (alive[i]) {
                set_data(rs[i], d, L, blob);
                rs[i].stack->update();
                x.setZero();
                rec[0] = rs[i].solver->solve(x) ? 1.0 : 0.0;
                for (int j = 0; j < d.n; ++j) rec[2 + j] = x[j];
            }
            try {
                RefStack cold;
                build(cold, d, L, blob, solvers::solver_back_ends::qpOASES);
                rec[1] = 1.0;
                final_x(cold, d, rec.data() + 2 + d.n);
            } catch (const std::exception&) {}
            fwrite(rec.data(), sizeof(double), rec.size(), o);
        }
    fclose(o); fclose(ob);
    return 0;
}

// The reference's nHQP front-end (src/solvers/nHQP.cpp) on instance blobs; see ref_nhqp.cpp.
//   ref_driver nhqp <desc.bin> <blobs.bin> <count> <min_sv_ratio> <ab_reg 0|1> <ns_reg 0|1> <out.bin>
// out.bin: int32 n, 0, count, 0 ; per instance: double ok, x[n]
extern "C" int nhqp_solve_blobs(const osb_stack_desc* d, const osb_layout* L, const double* blobs, int count,
                                double min_sv_ratio, int ab_regularisation, int ns_regularisation, double* out);
static int cmd_nhqp(int argc, char** argv)
{
    if (argc != 9) { fprintf(stderr, "nhqp: bad arguments\n"); return 2; }
    osb_stack_desc d; osb_layout L;
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(&d, sizeof(d), 1, f) != 1) { fprintf(stderr, "cannot read desc\n"); return 2; }
    fclose(f);
    if (osb_blob_layout(&d, &L)) return 2;
    const int count = atoi(argv[4]);
    std::vector<double> blobs((size_t)count * L.stride), out((size_t)count * (1 + d.n));
    f = fopen(argv[3], "rb");
    if (!f || fread(blobs.data(), sizeof(double), blobs.size(), f) != blobs.size()) return 2;
    fclose(f);
    nhqp_solve_blobs(&d, &L, blobs.data(), count, atof(argv[5]), atoi(argv[6]), atoi(argv[7]), out.data());
    FILE* o = fopen(argv[8], "wb");
    if (!o) return 2;
    int32_t hdr[4] = {d.n, 0, count, 0};
    fwrite(hdr, sizeof(int32_t), 4, o);
    fwrite(out.data(), sizeof(double), out.size(), o);
    fclose(o);
    return 0;
}

// CPU baseline: nproc worker PROCESSES (the reference has global statics and qpOASES a
// global message handler: threads are unsafe). Each worker owns
// count/nproc persistent solver objects built before the clock starts; per cycle the
// inputs drift a little (time-coherent, so qpOASES hot-starts as in production) and only
// AutoStack::update() + iHQP::solve() are timed with steady_clock.
static int cmd_bench(int argc, char** argv)
{
    if (argc != 9 && argc != 10) { fprintf(stderr, "bench: bad arguments\n"); return 2; }
    const int warm = (argc == 10) ? atoi(argv[9]) : 0;
    osb_stack_desc d; osb_layout L;
    if (synth_shape(argv[2], atof(argv[3]), &d) || osb_blob_layout(&d, &L)) return 2;
    const uint64_t seed = strtoull(argv[4], 0, 10);
    const int count = atoi(argv[5]);
    const double tight = atof(argv[6]);
    const int nproc = atoi(argv[7]), cycles = atoi(argv[8]);
    struct Shared { double busy_s; long solves; long failures; };
    Shared* sh = (Shared*)mmap(0, sizeof(Shared) * nproc, PROT_READ | PROT_WRITE,
                               MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    memset(sh, 0, sizeof(Shared) * nproc);
    auto wall0 = std::chrono::steady_clock::now();
    for (int p = 0; p < nproc; ++p) {
        pid_t pid = fork();
        if (pid == 0) {
            const int lo = (int)((long)count * p / nproc), hi = (int)((long)count * (p + 1) / nproc);
            std::vector<RefStack> rs(hi - lo);
            std::vector<double> blob(L.stride), base((size_t)(hi - lo) * L.stride);
            for (int i = lo; i < hi; ++i) {
                synth_fill(&d, &L, seed, (uint64_t)i, tight, base.data() + (size_t)(i - lo) * L.stride);
                build(rs[i - lo], d, L, base.data() + (size_t)(i - lo) * L.stride,
                      solvers::solver_back_ends::qpOASES);
            }
            double busy = 0; long solves = 0, fails = 0;
            Eigen::VectorXd x(d.n);
            synth_rng r; synth_seed(&r, seed ^ 0xabcdefULL, (uint64_t)p);
            for (int c = -warm; c < cycles; ++c)
                for (int i = lo; i < hi; ++i) {
                    const double* b0 = base.data() + (size_t)(i - lo) * L.stride;
                    memcpy(blob.data(), b0, sizeof(double) * L.stride);
                    for (int k = 0; k < d.n_levels; ++k) { // drift the task data only
                        if (d.task_kind[k] == OSB_TASK_DENSE)
                            for (int j = 0; j < d.m[k] * d.n; ++j)
                                blob[L.A[k] + j] += 0.01 * synth_normal(&r);
                        for (int j = 0; j < d.m[k]; ++j)
                            blob[L.b[k] + j] *= 1.0 + 0.01 * synth_normal(&r);
                    }
                    set_data(rs[i - lo], d, L, blob.data());
                    auto t0 = std::chrono::steady_clock::now();
                    rs[i - lo].stack->update();
                    bool ok = rs[i - lo].solver->solve(x);
                    auto t1 = std::chrono::steady_clock::now();
                    if (c < 0) continue; /* warm-up cycle: not timed */
                    busy += std::chrono::duration<double>(t1 - t0).count();
                    ++solves;
                    if (!ok) ++fails;
                }
            sh[p].busy_s = busy; sh[p].solves = solves; sh[p].failures = fails;
            _exit(0);
        }
    }
    int st;
    while (wait(&st) > 0) {}
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
    double max_busy = 0, sum_busy = 0; long solves = 0, fails = 0;
    for (int p = 0; p < nproc; ++p) {
        max_busy = std::max(max_busy, sh[p].busy_s); sum_busy += sh[p].busy_s;
        solves += sh[p].solves; fails += sh[p].failures;
    }
    // throughput = all solves / slowest worker's timed seconds (workers run concurrently)
    printf("{\"shape\": \"%s\", \"solves\": %ld, \"failures\": %ld, \"nproc\": %d, "
           "\"max_busy_s\": %.6f, \"sum_busy_s\": %.6f, \"wall_s\": %.3f, "
           "\"stacks_per_s\": %.3f, \"us_per_stack_per_core\": %.3f}\n",
           argv[2], solves, fails, nproc, max_busy, sum_busy, wall,
           solves / max_busy, 1e6 * sum_busy / solves);
    return 0;
}

// One BackEnd-level QP through the reference plugin boundary (BackEndFactory -> initProblem
// -> solve). in.bin: int32 n, nc ; double eps_factor, H[n*n] row-major, g[n], C[nc*n]
// row-major, lC[nc], uC[nc], l[n], u[n]. Prints "ok x0 x1 ...".
static int cmd_qp(int argc, char** argv)
{
    if (argc != 3) return 2;
    FILE* f = fopen(argv[2], "rb");
    int32_t hdr[2];
    if (!f || fread(hdr, sizeof(int32_t), 2, f) != 2) return 2;
    const int n = hdr[0], nc = hdr[1];
    std::vector<double> buf(1 + n * n + n + nc * n + 2 * nc + 2 * n);
    if (fread(buf.data(), sizeof(double), buf.size(), f) != buf.size()) return 2;
    fclose(f);
    const double* p = buf.data();
    const double eps_factor = *p++;
    Eigen::MatrixXd H = Eigen::Map<const RowMat>(p, n, n); p += n * n;
    Eigen::VectorXd g = Eigen::Map<const Eigen::VectorXd>(p, n); p += n;
    Eigen::MatrixXd C(nc, n);
    if (nc > 0) C = Eigen::Map<const RowMat>(p, nc, n);
    p += nc * n;
    Eigen::VectorXd lC = Eigen::Map<const Eigen::VectorXd>(p, nc); p += nc;
    Eigen::VectorXd uC = Eigen::Map<const Eigen::VectorXd>(p, nc); p += nc;
    Eigen::VectorXd l = Eigen::Map<const Eigen::VectorXd>(p, n); p += n;
    Eigen::VectorXd u = Eigen::Map<const Eigen::VectorXd>(p, n); p += n;
    if (nc == 0) C.resize(0, 0);
    solvers::BackEnd::Ptr be = solvers::BackEndFactory(solvers::solver_back_ends::qpOASES, n, nc,
                                                       HST_SEMIDEF, eps_factor);
    bool ok = be->initProblem(H, g, C, lC, uC, l, u);
    printf("%d", ok ? 1 : 0);
    for (int i = 0; i < n; ++i) printf(" %.17g", be->getSolution()[i]);
    printf("\n");
    return 0;
}

// ---- model-free known-answer tests of the reference, replayed through this build ------
static int check(const char* name, bool ok) { printf("%-52s %s\n", name, ok ? "PASS" : "FAIL"); return ok ? 0 : 1; }

static int cmd_kat()
{
    int bad = 0;
    using namespace OpenSoT::solvers;
    { // tests/solvers/TestQPOases.cpp:346-377 testSimpleProblem (H=I, g=[-5,5] -> x=-g, 1e-14)
        Eigen::MatrixXd H = Eigen::MatrixXd::Identity(2, 2), A(2, 2);
        Eigen::VectorXd g(2), l(2), u(2), lA(2), uA(2);
        g << -5.0, 5.0; A.setZero(); l << -10, -10; u << 10, 10; lA << -10, -10; uA << 10, 10;
        BackEnd::Ptr be = BackEndFactory(solver_back_ends::qpOASES, 2, 2, HST_IDENTITY, 1e-9);
        bool ok = be->initProblem(H, g, A, lA, uA, l, u) && be->solve();
        Eigen::VectorXd s = be->getSolution();
        ok = ok && std::fabs(s[0] - 5.0) < 1e-14 && std::fabs(s[1] + 5.0) < 1e-14;
        for (int i = 0; i < 1000 && ok; ++i) { ok = be->solve(); s = be->getSolution();
            ok = ok && std::fabs(s[0] - 5.0) < 1e-14 && std::fabs(s[1] + 5.0) < 1e-14; }
        bad += check("TestQPOases.testSimpleProblem (x1000)", ok);
        BackEnd::Ptr be1 = BackEndFactory(solver_back_ends::qpOASES, 2, 2, HST_IDENTITY, 1.0);
        bad += check("TestQPOases.cpp:833 eps == 2.221e-13",
                     std::fabs(be1->getEpsRegularisation() - 2.221e-13) < 1e-15);
    }
    { // tests/solvers/TestQPOases.cpp:208-272 test_update_constraint
        Eigen::MatrixXd T(1, 3); T << 1, 1, 1;
        Eigen::VectorXd b(1); b << 10;
        Eigen::MatrixXd A = Eigen::MatrixXd::Zero(1, 3);
        Eigen::VectorXd lA = Eigen::VectorXd::Zero(1), uA = Eigen::VectorXd::Zero(1), l(3), u(3);
        l.setConstant(-10); u.setConstant(10);
        BackEnd::Ptr qp = BackEndFactory(solver_back_ends::qpOASES, 3, 1, HST_SEMIDEF, 1e4);
        bool ok = qp->initProblem(T.transpose() * T, -1. * T.transpose() * b, A, lA, uA, l, u) && qp->solve();
        Eigen::VectorXd s = qp->getSolution();
        for (int i = 0; i < 3; ++i) ok = ok && std::fabs(s[i] - 3.333) < 1e-3;
        A << 1, 0, 1; lA << 20; uA = lA;
        ok = ok && qp->updateConstraints(A, lA, uA) && qp->solve();
        s = qp->getSolution();
        ok = ok && std::fabs(s[0] - 10.) < 1e-6 && std::fabs(s[1] + 10.) < 1e-6 && std::fabs(s[2] - 10.) < 1e-6;
        bad += check("TestQPOases.test_update_constraint", ok);
    }
    { // tests/solvers/TestQPOases.cpp:274-340 test_update_task
        BackEnd::Ptr qp = BackEndFactory(solver_back_ends::qpOASES, 3, 0, HST_UNKNOWN, 1.);
        Eigen::MatrixXd T(2, 3); T << 1, 1, 1, 0, 1, 1;
        Eigen::VectorXd b(2); b << 6, 5;
        Eigen::MatrixXd A(0, 0); Eigen::VectorXd lA, uA, l(3), u(3);
        l.setConstant(-10); u.setConstant(10);
        bool ok = qp->initProblem(T.transpose() * T, -1. * T.transpose() * b, A, lA, uA, l, u) && qp->solve();
        Eigen::VectorXd s = qp->getSolution();
        ok = ok && std::fabs(s[0] - 1.) < 1e-6 && std::fabs(s[1] - 2.5) < 1e-6 && std::fabs(s[2] - 2.5) < 1e-6;
        T.resize(3, 3); T << 1, 1, 1, 0, 1, 1, 1, 1, 0; b.resize(3); b << 6, 5, 3;
        ok = ok && qp->updateTask(T.transpose() * T, -1. * T.transpose() * b) && qp->solve();
        s = qp->getSolution();
        ok = ok && std::fabs(s[0] - 1.) < 1e-6 && std::fabs(s[1] - 2.) < 1e-6 && std::fabs(s[2] - 3.) < 1e-6;
        T.resize(4, 3); T << 1, 1, 1, 0, 1, 1, 1, 1, 0, 1, 0, 1; b.resize(4); b << 6, 5, 3, 3;
        ok = ok && qp->updateTask(T.transpose() * T, -1. * T.transpose() * b) && qp->solve();
        s = qp->getSolution();
        ok = ok && std::fabs(s[0] - .5714) < 1e-4 && std::fabs(s[1] - 2.5714) < 1e-4 && std::fabs(s[2] - 2.5714) < 1e-4;
        bad += check("TestQPOases.test_update_task", ok);
    }
    { // tests/solvers/TestiHQP.cpp:73-146 : H == A'A, g == -A'b, x solves Ax=b
        Eigen::MatrixXd A(2, 3); A << 1, 2, 3, 4, 5, 6;
        Eigen::VectorXd b(2); b << 1, 2;
        auto t = std::make_shared<tasks::GenericTask>("t", A, b);
        tasks::Aggregated::TaskPtr tp = t;
        AutoStack::Ptr st = std::make_shared<AutoStack>(tp);
        st->update();
        iHQP s(*st, 2e2, solver_back_ends::qpOASES);
        Eigen::VectorXd x(3);
        bool ok = s.solve(x);
        BackEnd::Ptr be; s.getBackEnd(0, be);
        Eigen::MatrixXd H = A.transpose() * A;
        ok = ok && (be->getH() - H).cwiseAbs().maxCoeff() < 1e-9;
        ok = ok && (be->getg() + A.transpose() * b).cwiseAbs().maxCoeff() < 1e-12;
        ok = ok && (A * x - b).cwiseAbs().maxCoeff() < 1e-8;
        bad += check("TestiHQP H=A'A g=-A'b Ax=b", ok);
    }
    printf("%s\n", bad ? "KAT FAILURES" : "ALL KATS PASS");
    return bad;
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: ref_driver solve|solveblob|bench|kat ...\n"); return 2; }
    std::cout.setstate(std::ios_base::failbit); // BackEndFactory chats on std::cout per object
    if (!strcmp(argv[1], "solve")) return cmd_solve(argc, argv);
    if (!strcmp(argv[1], "solveblob")) return cmd_solveblob(argc, argv);
    if (!strcmp(argv[1], "solveboth")) return cmd_solveboth(argc, argv);
    if (!strcmp(argv[1], "sequence")) return cmd_sequence(argc, argv);
    if (!strcmp(argv[1], "nhqp")) return cmd_nhqp(argc, argv);
    if (!strcmp(argv[1], "bench")) return cmd_bench(argc, argv);
    if (!strcmp(argv[1], "kat")) return cmd_kat();
    if (!strcmp(argv[1], "qp")) return cmd_qp(argc, argv);
    fprintf(stderr, "unknown command %s\n", argv[1]);
    return 2;
}

// test_dropin.cpp -- TEST: the same C++ OpenSoT stack objects solved three ways.
//   (1) reference: one iHQP + qpOASES per stack                 (CPU, the oracle)
//   (2) ours     : OpenSoT::b200::BatchediHQP over all stacks   (one GPU solve through the C ABI)
//   (3) plugin   : unmodified iHQP driving libOpenSotBackEndODYS.so = B200BackEnd, level by level
// The stacks are built with the reference's own Math-of-Tasks operators (oracle/ref_stack.h).
// Prints one JSON line; the pytest wrapper (tests/test_gpu_host_cpp.py) asserts on it.
// usage: test_dropin <shape> <count> <tight> <seed> <n_plugin>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <OpenSoT/constraints/Aggregated.h>
#include <iostream>
#include <random>

#include "BatchedHCOD.h"
#include "BatchediHQP.h"
#include "BatchednHQP.h"
#include "ref_stack.h"
#include "opensot_b200_synth.h"
#include "../../opensot_b200/csrc/osb_layout.c"

using namespace refstack;

// ---- every member of the iHQP / HCOD surface that BatchediHQP / BatchedHCOD mirror (iHQP.h:59-193,
// HCOD.h:33-41), each against the reference object built with the same arguments -----------------------
static double maxabs(const Eigen::MatrixXd& a) { return a.size() ? a.cwiseAbs().maxCoeff() : 0.0; }

static int surface(int argc, char** argv)
{
    if (argc != 6) { fprintf(stderr, "usage: test_dropin surface shape count tight seed\n"); return 2; }
    std::cout.setstate(std::ios_base::failbit);
    osb_stack_desc d; osb_layout L;
    if (synth_shape(argv[2], 2e2, &d) || osb_blob_layout(&d, &L)) return 2;
    const int count = atoi(argv[3]); const double tight = atof(argv[4]);
    const uint64_t seed = strtoull(argv[5], 0, 10);
    const int n = d.n;
    typedef OpenSoT::b200::BatchediHQP BQ;
    std::vector<double> blob(L.stride);
    std::vector<RefStack> rs(count);
    std::vector<BQ::Stack> stacks;
    std::vector<BQ::ConstraintPtr> bounds, globals;
    std::vector<OpenSoT::AutoStack::Ptr> autos;
    for (int i = 0; i < count; ++i) {
        synth_fill(&d, &L, seed, (uint64_t)i, tight, blob.data());
        build(rs[i], d, L, blob.data(), solvers::solver_back_ends::qpOASES, true);
        stacks.push_back(rs[i].stack->getStack());
        std::list<BQ::ConstraintPtr> bl(rs[i].bounds.begin(), rs[i].bounds.end());
        bounds.push_back(std::make_shared<constraints::Aggregated>(bl, n));
        bounds.back()->update();
        globals.push_back(rs[i].global_con);
        autos.push_back(rs[i].stack);
    }
    const bool has_global = d.c_in > 0 && d.c_eq == 0;
    std::string js = "{"; char buf[1024];
#define JS(...) do { snprintf(buf, sizeof(buf), __VA_ARGS__); js += buf; } while (0)
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 1\n");
    // (1) Stack& / bounds / globalConstraints constructors, single and batched, against iHQP with the same arguments
    {
        double e1 = 0, e2 = 0, e3 = 0, eb = 0;
        if (d.c_in == 0 && d.c_eq == 0) {
            Eigen::MatrixXd Xb; BQ batched(stacks, bounds, 2e2, 0);
            bool ok = batched.solve(Xb);
            for (int i = 0; i < std::min(count, 6); ++i) {
                solvers::iHQP ref(stacks[i], bounds[i], 2e2, solvers::solver_back_ends::qpOASES);
                Eigen::VectorXd xr(n), x1(n); ref.solve(xr);
                BQ one(stacks[i], bounds[i], 2e2, solvers::solver_back_ends::qpOASES);
                ok = one.solve(x1) && ok;
                e2 = std::max(e2, maxabs(x1 - xr)); eb = std::max(eb, maxabs(Xb.col(i) - xr));
                solvers::iHQP refn(stacks[i], 2e2, solvers::solver_back_ends::qpOASES);    // no bounds at all
                BQ onen(stacks[i], 2e2, std::vector<solvers::solver_back_ends>(d.n_levels, solvers::solver_back_ends::qpOASES));
                Eigen::VectorXd xn(n), xrn(n); refn.solve(xrn); ok = onen.solve(xn) && ok;
                e1 = std::max(e1, maxabs(xn - xrn));
            }
            JS("\"ctor_stack_err\": %.3e, \"ctor_stack_bounds_err\": %.3e, \"ctor_batched_bounds_err\": %.3e, \"ctor_ok\": %d, ", e1, e2, eb, ok ? 1 : 0);
        }
        if (has_global) {
            Eigen::MatrixXd Xb; BQ batched(stacks, bounds, globals, 2e2, 0);
            bool ok = batched.solve(Xb);
            for (int i = 0; i < std::min(count, 6); ++i) {
                solvers::iHQP ref(stacks[i], bounds[i], globals[i], 2e2, solvers::solver_back_ends::qpOASES);
                Eigen::VectorXd xr(n), x1(n); ref.solve(xr);
                BQ one(stacks[i], bounds[i], globals[i], 2e2);
                ok = one.solve(x1) && ok;
                e3 = std::max(e3, maxabs(x1 - xr)); eb = std::max(eb, maxabs(Xb.col(i) - xr));
            }
            JS("\"ctor_global_err\": %.3e, \"ctor_batched_global_err\": %.3e, \"ctor_global_ok\": %d, ", e3, eb, ok ? 1 : 0);
        }
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 2\n");
    // (2) through the Solver<> base (Solver.h:84) and the iHQP(AutoStack&) signature
    {
        BQ::Base::SolverPtr base = std::make_shared<BQ>(*rs[0].stack, 2e2, solvers::solver_back_ends::qpOASES);
        Eigen::VectorXd x(n), xr(n);
        rs[0].stack->update(); rs[0].solver->solve(xr);
        const bool ok = base->solve(x);
        base->setSolverID("b200");
        JS("\"solver_base_err\": %.3e, \"solver_base_ok\": %d, ", maxabs(x - xr), (ok && base->getSolverID() == "b200") ? 1 : 0);
    }
    BQ solver(autos, 2e2, 0);
    Eigen::MatrixXd X, Xref(n, count);
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 3\n");
    // (3) getObjective / getBackEnd against the reference's back-ends
    {
        bool ok = solver.updateAndSolve(X);
        double eobj = 0, esol = 0, eH = 0, eg = 0, eA = 0; int rows_ok = 1;
        for (int i = 0; i < count; ++i) {
            Eigen::VectorXd x(n); rs[i].stack->update(); rs[i].solver->solve(x); Xref.col(i) = x;
            if (i >= 6) continue;
            for (int k = 0; k < d.n_levels; ++k) {
                double vr = 0, vo = 0;
                rs[i].solver->getObjective(k, vr); ok = solver.getObjective(k, vo, i) && ok;
                eobj = std::max(eobj, std::fabs(vr - vo) / std::max(1e-3, std::fabs(vr)));
                solvers::BackEnd::Ptr br, bo;
                rs[i].solver->getBackEnd(k, br); ok = solver.getBackEnd(k, bo, i) && ok;
                // the last level fixes all of x; an earlier level fixes A_k x, the rest of its x is held by eps alone
                // (there qpOASES itself is 1e-7..1e-5 off the exact minimiser)
                if (k == d.n_levels - 1) esol = std::max(esol, maxabs(br->getSolution() - bo->getSolution()));
                else {
                    const Eigen::MatrixXd& Ak = rs[i].stack->getStack()[k]->getA();
                    esol = std::max(esol, maxabs(Ak * (br->getSolution() - bo->getSolution())));
                }
                Eigen::MatrixXd Hr = br->getH(); Hr.diagonal().array() -= br->getEpsRegularisation();   // the reference leaves eps on _H
                eH = std::max(eH, maxabs(Hr - bo->getH())); eg = std::max(eg, maxabs(br->getg() - bo->getg()));
                if (br->getA().rows() != bo->getA().rows()) { rows_ok = 0; fprintf(stderr, "level %d rows ref %d ours %d\n", k, (int)br->getA().rows(), (int)bo->getA().rows()); }
                else if (br->getA().rows()) {
                    eA = std::max(eA, maxabs(br->getA() - bo->getA()));
                    for (int r = 0; r < br->getlA().size(); ++r) {      // finite sides only (the reference clips to +-1e20 in place)
                        if (std::fabs(br->getlA()[r]) < 1e19) eA = std::max(eA, std::fabs(br->getlA()[r] - bo->getlA()[r]));
                        if (std::fabs(br->getuA()[r]) < 1e19) eA = std::max(eA, std::fabs(br->getuA()[r] - bo->getuA()[r]));
                    }
                }
                if (std::fabs(bo->getEpsRegularisation() - br->getEpsRegularisation()) > 1e-18) { rows_ok = 0; fprintf(stderr, "eps ref %g ours %g\n", br->getEpsRegularisation(), bo->getEpsRegularisation()); }
            }
        }
        Eigen::MatrixXd X1; solver.getLevelSolutions(d.n_levels - 1, X1);
        JS("\"solve_err\": %.3e, \"objective_rel_err\": %.3e, \"backend_solution_err\": %.3e, \"backend_H_err\": %.3e, "
               "\"backend_g_err\": %.3e, \"backend_A_err\": %.3e, \"backend_rows_ok\": %d, \"last_level_is_x\": %d, \"surface_ok\": %d, ",
               maxabs(X - Xref), eobj, esol, eH, eg, eA, rows_ok, maxabs(X1 - X) == 0.0 ? 1 : 0, ok ? 1 : 0);
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 3b\n");
    // (3b) Solver<>::log (Solver.h:105-115) -> iHQP::_log (iHQP.cpp:403-415) -> BackEnd::log of every level (BackEnd.cpp:111-130):
    //      the same variable names in the same order with the same sizes and values as the reference solver of that instance
    {
        auto lr = std::make_shared<XBot::MatLogger2>(), lo = std::make_shared<XBot::MatLogger2>();
        const int inst = std::min(count - 1, 2);
        rs[inst].solver->setSolverID("wb"); rs[inst].solver->log(lr);
        solver.setSolverID("wb"); solver.setLogInstance(inst); solver.log(lo);
        int names_ok = lr->entries.size() == lo->entries.size() && !lr->entries.empty();
        double e = 0;
        for (size_t k = 0; names_ok && k < lr->entries.size(); ++k) {
            const auto &a = lr->entries[k], &b = lo->entries[k];
            if (a.name != b.name || a.size != b.size) { names_ok = 0; fprintf(stderr, "log entry %zu: ref %s[%ld] ours %s[%ld]\n", k, a.name.c_str(), a.size, b.name.c_str(), b.size); break; }
            // an earlier level's own solution is held by eps alone off its task's row space (see (3)): compared loosely
            const bool loose = a.name.find("solution_") != std::string::npos && a.name != "wb_solution_" + std::to_string(d.n_levels - 1);
            const double err = std::fabs(a.checksum - b.checksum) / std::max(1.0, std::fabs(a.checksum));
            if (!loose) e = std::max(e, err);
            else if (err > 1e-2) names_ok = 0;
        }
        if (lr->entries.size() != lo->entries.size()) fprintf(stderr, "log entries: ref %zu ours %zu\n", lr->entries.size(), lo->entries.size());
        JS("\"log_names_ok\": %d, \"log_entries\": %d, \"log_checksum_err\": %.3e, ", names_ok, (int)lr->entries.size(), e);
        rs[inst].solver->setSolverID(""); solver.setSolverID("");
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 4\n");
    // (4) setOptions / getOptions (boost::any holding the engine's option struct), out-of-range levels
    {
        boost::any a; bool ok = solver.getOptions(0, a);
        osb_options o = boost::any_cast<osb_options>(a);
        o.max_iter = 1;
        ok = solver.setOptions(0, boost::any(o)) && ok;
        Eigen::MatrixXd Xm; const bool solved = solver.solve(Xm);      // one working-set change per level cannot be enough
        int maxit = 0; for (int s : solver.getStatus()) maxit += ((s & 0xff) == OSB_MAX_ITER);
        o.max_iter = 0; ok = solver.setOptions(d.n_levels - 1, boost::any(o)) && ok;
        const bool back = solver.solve(Xm) && maxabs(Xm - Xref) <= 1e-6;
        const bool range = !solver.setOptions(d.n_levels, boost::any(o)) && !solver.getOptions(d.n_levels, a) &&
                           !solver.setOptions(0, boost::any(std::string("not options"))) && solver.getBackEndName(0) == "B200" &&
                           solver.getBackEndName(d.n_levels) == "" && (int)solver.getNumberOfTasks() == d.n_levels;
        JS("\"options_ok\": %d, \"max_iter_hits\": %d, \"max_iter_solve_false\": %d, \"options_restored\": %d, \"range_checks\": %d, ",
               ok ? 1 : 0, maxit, solved ? 0 : 1, back ? 1 : 0, range ? 1 : 0);
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 5\n");
    // (5) setEpsRegularisation(eps) and (eps, i): absolute values, like QPOasesBackEnd.cpp:358-372
    {
        double e_all = 0, e_one = 0; bool ok = true;
        ok = solver.setEpsRegularisation(1e-3) && ok;
        for (int i = 0; i < count; ++i) rs[i].solver->setEpsRegularisation(1e-3);
        Eigen::MatrixXd Xe; ok = solver.solve(Xe) && ok;
        const double moved_all = maxabs(Xe - Xref);      // the shift must matter, or the comparison below says nothing
        for (int i = 0; i < count; ++i) { Eigen::VectorXd x(n); rs[i].solver->solve(x); e_all = std::max(e_all, maxabs(Xe.col(i) - x)); }
        const unsigned lvl = d.n_levels - 1;
        ok = solver.setEpsRegularisation(1e-2, lvl) && ok;
        for (int i = 0; i < count; ++i) rs[i].solver->setEpsRegularisation(1e-2, lvl);
        const Eigen::MatrixXd Xall = Xe;
        ok = solver.solve(Xe) && ok;
        const double moved_one = maxabs(Xe - Xall);
        for (int i = 0; i < count; ++i) { Eigen::VectorXd x(n); rs[i].solver->solve(x); e_one = std::max(e_one, maxabs(Xe.col(i) - x)); }
        const bool neg = !solver.setEpsRegularisation(-1.0) && !solver.setEpsRegularisation(1.0, d.n_levels);
        const double def = 2e2 * OSB_QPOASES_BASE_EPS;
        solver.setEpsRegularisation(def);
        for (int i = 0; i < count; ++i) rs[i].solver->setEpsRegularisation(def);
        JS("\"eps_all_err\": %.3e, \"eps_level_err\": %.3e, \"eps_all_moved\": %.3e, \"eps_level_moved\": %.3e, \"eps_ok\": %d, \"eps_rejects\": %d, ", e_all, e_one, moved_all, moved_one, ok ? 1 : 0, neg ? 1 : 0);
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 6\n");
    // (6) a nested constraints::Aggregated that carries BOTH a box and rows, passed with <<: the reference
    //     intersects the box of every member (constraints/Aggregated.cpp:118-148)
    {
        double e = 0; int active = 0; bool ok = true;
        std::vector<RefStack> ms(std::min(count, 8));
        std::vector<OpenSoT::AutoStack::Ptr> mstacks;
        std::vector<std::shared_ptr<solvers::iHQP>> mref;
        for (size_t i = 0; i < ms.size(); ++i) {
            osb_stack_desc d0 = d; d0.n_bound_sets = 0; osb_layout L0; osb_blob_layout(&d0, &L0);
            std::vector<double> b0(L0.stride);
            synth_fill(&d0, &L0, seed + 7, (uint64_t)i, tight, b0.data());
            build(ms[i], d0, L0, b0.data(), solvers::solver_back_ends::qpOASES, false);
            Eigen::VectorXd ub = Eigen::VectorXd::Constant(n, 0.004), lb = -ub;    // tight box: it must bind
            auto box = std::make_shared<constraints::GenericConstraint>("box", ub, lb, n);
            Eigen::MatrixXd C = Eigen::MatrixXd::Zero(2, n); C(0, 0) = 1; C(0, 1) = 1; C(1, 2) = 1; C(1, 3) = -1;
            AffineHelper var(C, Eigen::VectorXd::Zero(2));
            auto rows = std::make_shared<constraints::GenericConstraint>("rows", var, Eigen::VectorXd::Constant(2, 0.003),
                                                                         Eigen::VectorXd::Constant(2, -0.003),
                                                                         constraints::GenericConstraint::Type::CONSTRAINT);
            std::list<BQ::ConstraintPtr> both = {box, rows};
            auto mixed = std::make_shared<constraints::Aggregated>(both, n);
            ms[i].stack = ms[i].stack << mixed;
            ms[i].stack->update();
            mstacks.push_back(ms[i].stack);
            mref.push_back(std::make_shared<solvers::iHQP>(*ms[i].stack, 2e2, solvers::solver_back_ends::qpOASES));
        }
        BQ msolver(mstacks, 2e2, 0);
        Eigen::MatrixXd Xm; ok = msolver.updateAndSolve(Xm);
        for (size_t i = 0; i < ms.size(); ++i) {
            Eigen::VectorXd x(n); ms[i].stack->update(); mref[i]->solve(x);
            if (getenv("OSB_TEST_TRACE")) {     // arbitration of a disagreement: a cold qpOASES and eiQuadProg on the same stack
                double dc = -1, de = -1;
                try { solvers::iHQP cold(*ms[i].stack, 2e2, solvers::solver_back_ends::qpOASES); Eigen::VectorXd xc(n); if (cold.solve(xc)) dc = maxabs(Xm.col(i) - xc); } catch (...) {}
                try { solvers::iHQP ei(*ms[i].stack, 2e2, solvers::solver_back_ends::eiQuadProg); Eigen::VectorXd xe(n); if (ei.solve(xe)) de = maxabs(Xm.col(i) - xe); } catch (...) {}
                fprintf(stderr, "\nmixed %zu: |ours-hot| %.2e |ours-cold| %.2e |ours-eiquadprog| %.2e (-1: that solver failed)\n", i, maxabs(Xm.col(i) - x), dc, de);
                // which of the two is the better lexicographic answer: weighted residual of every level, worst violation of the box and the rows
                const Eigen::VectorXd xo = Xm.col(i);
                for (auto& tk : ms[i].stack->getStack()) {
                    const Eigen::VectorXd ro = tk->getA() * xo - tk->getb(), rr = tk->getA() * x - tk->getb();
                    fprintf(stderr, "   level cost ours %.12e reference %.12e\n", ro.dot(tk->getWeight() * ro), rr.dot(tk->getWeight() * rr));
                }
                fprintf(stderr, "   box violation ours %.2e reference %.2e, |x|^2 ours %.6e reference %.6e\n",
                        std::max(0.0, xo.cwiseAbs().maxCoeff() - 0.004), std::max(0.0, x.cwiseAbs().maxCoeff() - 0.004), xo.squaredNorm(), x.squaredNorm());
            }
            e = std::max(e, maxabs(Xm.col(i) - x));
            for (int j = 0; j < n; ++j) active += (std::fabs(std::fabs(Xm(j, i)) - 0.004) < 1e-9);
        }
        JS("\"mixed_aggregated_err\": %.3e, \"mixed_box_active\": %d, \"mixed_ok\": %d, ", e, active, ok ? 1 : 0);
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 7\n");
    // (7) an equality row that stops being one after construction: the engine is rebuilt, not failed
    if (d.c_eq > 0) {
        double e = 0; bool ok = true;
        for (int i = 0; i < count; ++i) {       // feet: the equality becomes a band of +-1e-3
            Eigen::VectorXd de = rs[i].feet->getb();
            rs[i].feet->setb(de);
        }
        // TaskToConstraint keeps lA == uA; widen through a second stack object instead
        std::vector<RefStack> ws(std::min(count, 8));
        std::vector<OpenSoT::AutoStack::Ptr> wstacks;
        std::vector<std::shared_ptr<solvers::iHQP>> wref;
        std::vector<constraints::GenericConstraint::Ptr> band;
        Eigen::MatrixXd C = Eigen::MatrixXd::Zero(2, n); C(0, 0) = 1; C(0, 4) = 1; C(1, 5) = 1; C(1, 6) = 1;
        for (size_t i = 0; i < ws.size(); ++i) {
            synth_fill(&d, &L, seed + 9, (uint64_t)i, tight, blob.data());
            build(ws[i], d, L, blob.data(), solvers::solver_back_ends::qpOASES, false);
            AffineHelper var(C, Eigen::VectorXd::Zero(2));
            band.push_back(std::make_shared<constraints::GenericConstraint>("band", var, Eigen::VectorXd::Zero(2), Eigen::VectorXd::Zero(2),
                                                                            constraints::GenericConstraint::Type::CONSTRAINT));
            ws[i].stack = ws[i].stack << band.back();
            ws[i].stack->update();
            wstacks.push_back(ws[i].stack);
            wref.push_back(std::make_shared<solvers::iHQP>(*ws[i].stack, 2e2, solvers::solver_back_ends::qpOASES));
        }
        BQ wsolver(wstacks, 2e2, 0);
        Eigen::MatrixXd Xw; ok = wsolver.updateAndSolve(Xw);
        const int eq_before = wsolver.getDescriptor().c_eq;
        for (size_t i = 0; i < ws.size(); ++i) {
            AffineHelper var(C, Eigen::VectorXd::Zero(2));
            band[i]->setConstraint(var, Eigen::VectorXd::Constant(2, 0.002), Eigen::VectorXd::Constant(2, -0.001));
        }
        ok = wsolver.updateAndSolve(Xw) && ok;
        for (size_t i = 0; i < ws.size(); ++i) {
            Eigen::VectorXd x(n); ws[i].stack->update(); wref[i]->solve(x);
            e = std::max(e, maxabs(Xw.col(i) - x));
        }
        JS("\"eq_to_ineq_err\": %.3e, \"eq_rows_before\": %d, \"eq_rows_after\": %d, \"eq_to_ineq_ok\": %d, ", e, eq_before,
               wsolver.getDescriptor().c_eq, ok ? 1 : 0);
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 9\n");
    // (9) an identity task matrix (engine: OSB_TASK_IDENTITY, nothing stored) that stops being one: rebuilt as dense
    if (d.task_kind[d.n_levels - 1] == OSB_TASK_IDENTITY) {
        const int last = d.n_levels - 1;
        const int kind_before = solver.getDescriptor().task_kind[last];
        for (int i = 0; i < count; ++i) {
            Eigen::MatrixXd A = Eigen::MatrixXd::Identity(n, n);
            A(0, 1) = 0.25; A(3, 2) = -0.5 + 0.01 * i;
            rs[i].level_task[last]->setA(A);
        }
        Eigen::MatrixXd Xi; const bool ok = solver.updateAndSolve(Xi);
        double e = 0;
        for (int i = 0; i < count; ++i) { Eigen::VectorXd x(n); rs[i].stack->update(); rs[i].solver->solve(x); e = std::max(e, maxabs(Xi.col(i) - x)); }
        JS("\"identity_to_dense_err\": %.3e, \"identity_kind_before\": %d, \"identity_kind_after\": %d, \"identity_to_dense_ok\": %d, ",
           e, kind_before, solver.getDescriptor().task_kind[last], ok ? 1 : 0);
        for (int i = 0; i < count; ++i) { rs[i].level_task[last]->setA(Eigen::MatrixXd::Identity(n, n)); rs[i].stack->update(); }
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 11\n");
    // (11) a weight that gains off-diagonal terms and a linear cost term that appears after construction (Task::setWeight, GenericTask::setc):
    //      the reference just uses them (iHQP.cpp:129-162); the engine is rebuilt for the wider class instead of failing. For the two to
    //      matter, level 1 asks for more than the box allows (its residual is then weighed by W) and c lies in the row space of level 0
    //      (it then shifts that level's target; any other c only pushes null-space directions into the box).
    if (d.n_levels > 1 && d.task_kind[0] == OSB_TASK_DENSE) {
        const int wk_before = solver.getDescriptor().weight_kind[1], c_before = solver.getDescriptor().has_c[0];
        const int m1 = d.m[1];
        std::vector<Eigen::MatrixXd> W_old(count);
        std::vector<Eigen::VectorXd> b_old(count);
        Eigen::MatrixXd Xbase(n, count);
        for (int i = 0; i < count; ++i) {
            W_old[i] = rs[i].level_task[1]->getWeight(); b_old[i] = rs[i].level_task[1]->getb();
            rs[i].level_task[1]->setb(8.0 * b_old[i]);
            Eigen::VectorXd x(n); rs[i].stack->update(); rs[i].solver->solve(x); Xbase.col(i) = x;
        }
        for (int i = 0; i < count; ++i) {
            Eigen::MatrixXd W = W_old[i];
            for (int r = 0; r + 1 < m1; ++r) { W(r, r + 1) += 0.3 * std::sqrt(W(r, r) * W(r + 1, r + 1)); W(r + 1, r) = W(r, r + 1); }   // stays positive definite
            rs[i].level_task[1]->setWeight(W);
            Eigen::VectorXd v = Eigen::VectorXd::Constant(d.m[0], 2e-3); v[0] = -1e-3 - 1e-4 * i;
            rs[i].level_task[0]->setc(rs[i].level_task[0]->getA().transpose() * v);
        }
        Eigen::MatrixXd Xw; const bool ok_all = solver.updateAndSolve(Xw);
        double e = 0, moved = 0;
        bool agree = true, ref_all = true;      // an over-asked level can make a later level's own rows inconsistent (C35h): both must say so
        for (int i = 0; i < count; ++i) {
            Eigen::VectorXd x(n); rs[i].stack->update(); const bool rok = rs[i].solver->solve(x);
            ref_all = ref_all && rok;
            if (rok != (solver.getStatus()[i] == OSB_OK)) { agree = false; fprintf(stderr, "weight_c_drift: instance %d reference solved %d, ours status %d\n", i, (int)rok, solver.getStatus()[i]); }
            if (!rok) continue;
            e = std::max(e, maxabs(Xw.col(i) - x)); moved = std::max(moved, maxabs(x - Xbase.col(i)));
        }
        const bool ok = agree && (ok_all == ref_all);
        JS("\"weight_c_drift_err\": %.3e, \"weight_c_drift_moved\": %.3e, \"weight_kind_before\": %d, \"weight_kind_after\": %d, \"has_c_before\": %d, \"has_c_after\": %d, \"weight_c_drift_ok\": %d, ",
           e, moved, wk_before, solver.getDescriptor().weight_kind[1], c_before, solver.getDescriptor().has_c[0], ok ? 1 : 0);
        for (int i = 0; i < count; ++i) {
            rs[i].level_task[1]->setWeight(W_old[i]); rs[i].level_task[1]->setb(b_old[i]);
            rs[i].level_task[0]->setc(Eigen::VectorXd::Zero(n)); rs[i].stack->update();
        }
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 10\n");
    // (10) constraints whose ROW COUNT differs between instances and changes from cycle to cycle (ConvexHull.cpp:47-52: one row
    //      per hull edge; the reference rebuilds its QP object, QPOasesBackEnd.cpp:229-244): a global one (`stack << hull`) and a
    //      task-local one (`task << hull`, that level only). BatchediHQP gives every constraint row SLOTS (null rows behind the
    //      last real one) and rebuilds its engine only when a constraint outgrows them.
    {
        double e = 0; bool ok = true; int binding = 0, max_rows = 0, min_rows = 1 << 20;
        const int cnt = std::min(count, 8), cycles = 7;
        std::vector<RefStack> hs(cnt);
        std::vector<OpenSoT::AutoStack::Ptr> hstacks;
        std::vector<std::shared_ptr<solvers::iHQP>> href;
        std::vector<constraints::GenericConstraint::Ptr> ghull(cnt), lhull(cnt);
        std::mt19937_64 gen(seed + 1234);
        std::normal_distribution<double> N01(0.0, 1.0);
        auto hull_rows = [&](constraints::GenericConstraint::Ptr& c, const char* name, const int rows) {
            Eigen::MatrixXd C(rows, n);
            Eigen::VectorXd ub(rows);
            for (int r = 0; r < rows; ++r) {
                for (int j = 0; j < n; ++j) C(r, j) = N01(gen);
                ub[r] = 0.0005 + 0.01 * std::fabs(N01(gen));         // 0 stays inside; tight enough to bind now and then
            }
            AffineHelper var(C, Eigen::VectorXd::Zero(rows));
            const Eigen::VectorXd lb = Eigen::VectorXd::Constant(rows, -1.0e20);
            if (!c) c = std::make_shared<constraints::GenericConstraint>(name, var, ub, lb, constraints::GenericConstraint::Type::CONSTRAINT);
            else c->setConstraint(var, ub, lb);
        };
        // rows per cycle: different on every instance, shrinking, then outgrowing everything seen before (cycles 3 and 5)
        auto g_rows = [&](int i, int c) { return c == 3 ? 9 + (i & 1) : (c == 5 ? 12 : 2 + (i + c) % 4); };
        auto l_rows = [&](int i, int c) { return c == 5 ? 6 : 1 + (i + 2 * c) % 3; };
        for (int i = 0; i < cnt; ++i) {
            synth_fill(&d, &L, seed + 11, (uint64_t)i, tight, blob.data());
            build(hs[i], d, L, blob.data(), solvers::solver_back_ends::qpOASES, false);
            hull_rows(ghull[i], "ghull", g_rows(i, 0));
            hull_rows(lhull[i], "lhull", l_rows(i, 0));
            tasks::Aggregated::TaskPtr t1 = hs[i].level_task[d.n_levels > 1 ? 1 : 0];
            t1 << lhull[i];
            hs[i].stack = hs[i].stack << ghull[i];
            hs[i].stack->update();
            hstacks.push_back(hs[i].stack);
            // (qpOASES may refuse to initialise a small, heavily constrained stack: then there is no reference for this scenario)
            try { href.push_back(std::make_shared<solvers::iHQP>(*hs[i].stack, 2e2, solvers::solver_back_ends::qpOASES)); }
            catch (const std::runtime_error&) { href.clear(); break; }
        }
        if ((int)href.size() == cnt) {
        BQ hsolver(hstacks, 2e2, 0);
        hsolver.setRowHeadroom(1);
        const int rows0 = hsolver.getDescriptor().c_in;
        int rebuilds_at[8] = {0};
        for (int c = 0; c < cycles; ++c) {
            for (int i = 0; i < cnt; ++i) {
                if (c > 0) { hull_rows(ghull[i], "ghull", g_rows(i, c)); hull_rows(lhull[i], "lhull", l_rows(i, c)); }
                max_rows = std::max(max_rows, g_rows(i, c)); min_rows = std::min(min_rows, g_rows(i, c));
            }
            Eigen::MatrixXd Xh; const bool all_ok = hsolver.updateAndSolve(Xh);
            if (!all_ok && !hsolver.getLastError().empty()) { ok = false; fprintf(stderr, "rows_change cycle %d: %s\n", c, hsolver.getLastError().c_str()); }
            rebuilds_at[c] = hsolver.getRebuilds();
            bool ref_all = true;
            for (int i = 0; i < cnt; ++i) {
                Eigen::VectorXd x(n); hs[i].stack->update();
                const bool rok = href[i]->solve(x);     // a stack with this many random half-spaces can be infeasible: both must say so
                ref_all = ref_all && rok;
                if (rok != (hsolver.getStatus()[i] == OSB_OK)) { ok = false; fprintf(stderr, "rows_change: cycle %d instance %d reference solved %d, ours status %d\n", c, i, (int)rok, hsolver.getStatus()[i]); }
                if (!rok) continue;
                e = std::max(e, maxabs(Xh.col(i) - x));
                const Eigen::VectorXd gx = ghull[i]->getAineq() * x;
                for (int r = 0; r < gx.size(); ++r) binding += (std::fabs(gx[r] - ghull[i]->getbUpperBound()[r]) < 1e-9);
            }
            // the BackEnd view of an instance reports the rows it has now, not the slots (iHQP.cpp:427-436)
            solvers::BackEnd::Ptr br, bo;
            const unsigned lvl = d.n_levels > 1 ? 1 : 0;
            href[0]->getBackEnd(lvl, br); ok = hsolver.getBackEnd(lvl, bo, 0) && ok;
            if (hsolver.getStatus()[0] == OSB_OK && (br->getA().rows() != bo->getA().rows() || maxabs(br->getA() - bo->getA()) > 1e-12)) { ok = false; fprintf(stderr, "rows_change: view rows ref %d ours %d\n", (int)br->getA().rows(), (int)bo->getA().rows()); }
        }
        // slots: the engine was rebuilt when rows outgrew them (cycles 3 and 5) and at no other cycle
        const int grew_only_when_needed = rebuilds_at[0] == 0 && rebuilds_at[2] == 0 && rebuilds_at[3] == 1 && rebuilds_at[4] == 1 &&
                                          rebuilds_at[5] == 2 && rebuilds_at[6] == 2;
        JS("\"rows_change_err\": %.3e, \"rows_change_ok\": %d, \"rows_change_binding\": %d, \"rows_change_rebuilds\": %d, "
           "\"rows_change_lazy_ok\": %d, \"rows_change_slots0\": %d, \"rows_change_slots\": %d, \"rows_change_min\": %d, \"rows_change_max\": %d, ",
           e, ok ? 1 : 0, binding, hsolver.getRebuilds(), grew_only_when_needed, rows0 - d.c_in, hsolver.getDescriptor().c_in - d.c_in, min_rows, max_rows);
        } else fprintf(stderr, "rows_change: the reference could not initialise these stacks, scenario skipped\n");
    }
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 8\n");
    // (8) HCOD(Stack&, bounds, damping) and HCOD(AutoStack&, damping) (HCOD.h:33-41), setDamping
    if (d.c_in == 0 && d.c_eq == 0) {
        OpenSoT::b200::BatchedHCOD h1(stacks[0], bounds[0], 0.0), h2(*rs[0].stack, 0.0);
        Eigen::VectorXd x1(n), x2(n);
        bool ok = h1.solve(x1) && h2.solve(x2);
        h1.setDamping(1e-3);
        Eigen::VectorXd x3(n); ok = h1.solve(x3) && ok;
        rs[0].solver->setEpsRegularisation(1e-6);
        Eigen::VectorXd xr(n); rs[0].solver->solve(xr);
        rs[0].solver->setEpsRegularisation(2e2 * OSB_QPOASES_BASE_EPS);
        JS("\"hcod_stack_bounds_err\": %.3e, \"hcod_autostack_err\": %.3e, \"hcod_damping_err\": %.3e, \"hcod_ok\": %d, ",
               maxabs(x1 - Xref.col(0)), maxabs(x2 - Xref.col(0)), maxabs(x3 - xr), (ok && h1.getDamping() == 1e-3) ? 1 : 0);
    }
    JS("\"shape\": \"%s\", \"count\": %d}", argv[2], count);
    printf("\n%s\n", js.c_str());
#undef JS
    return 0;
}

// ---- nHQP semantics: BatchednHQP against the reference's own nHQP (oracle/ref_nhqp.cpp drives the
// unmodified src/solvers/nHQP.cpp) on the same blobs, default settings and each switch flipped ---------------------
extern "C" int nhqp_solve_blobs(const osb_stack_desc* d, const osb_layout* L, const double* blobs, int count,
                                double min_sv_ratio, int ab_regularisation, int ns_regularisation, double* out);
static int nhqp(int argc, char** argv)
{
    if (argc != 6) { fprintf(stderr, "usage: test_dropin nhqp shape count tight seed\n"); return 2; }
    std::cout.setstate(std::ios_base::failbit);
    osb_stack_desc d; osb_layout L;
    if (synth_shape(argv[2], 2e2, &d) || osb_blob_layout(&d, &L)) return 2;
    const int count = atoi(argv[3]); const double tight = atof(argv[4]);
    const uint64_t seed = strtoull(argv[5], 0, 10);
    const int n = d.n;
    typedef OpenSoT::b200::BatchednHQP BN;
    std::vector<double> blobs((size_t)count * L.stride);
    std::vector<RefStack> rs(count);
    std::vector<BN::Stack> stacks;
    std::vector<BN::ConstraintPtr> bounds;
    for (int i = 0; i < count; ++i) {
        double* blob = blobs.data() + (size_t)i * L.stride;
        synth_fill(&d, &L, seed, (uint64_t)i, tight, blob);
        // a nearly dependent task row on the second level: a singular value below min_sv_ratio * sv_max, so that
        // the A/b damping of nHQP.cpp:236-280 is at work (Gaussian rows alone never trigger it)
        if (d.n_levels > 1 && d.task_kind[1] == OSB_TASK_DENSE && d.m[1] > 1)
            for (int j = 0; j < n; ++j) blob[L.A[1] + j] = blob[L.A[1] + n + j] + 1e-2 * blob[L.A[1] + j];
        build(rs[i], d, L, blob, solvers::solver_back_ends::qpOASES, false);
        stacks.push_back(rs[i].stack->getStack());
        std::list<BN::ConstraintPtr> all(rs[i].bounds.begin(), rs[i].bounds.end());
        if (rs[i].global_con) all.push_back(rs[i].global_con);
        bounds.push_back(std::make_shared<constraints::Aggregated>(all, n));
        bounds.back()->update();
    }
    std::string js = "{"; char buf[512];
    std::vector<double> ref((size_t)count * (1 + n));
    BN solver(stacks, bounds, 2e2, 0);
    Eigen::MatrixXd Xdef;
    struct Cfg { const char* name; double sv; int ab, ns; } cfgs[] = {{"default", 0.05, 1, 1}, {"no_ab", 0.05, 0, 1}, {"no_ns", 0.05, 1, 0}, {"sv20", 0.2, 1, 1}};
    for (const Cfg& c : cfgs) {
        nhqp_solve_blobs(&d, &L, blobs.data(), count, c.sv, c.ab, c.ns, ref.data());
        solver.setMinSingularValueRatio(c.sv);
        solver.setPerformAbRegularization(c.ab != 0);
        solver.setPerformSelectiveNullSpaceRegularization(c.ns != 0);
        Eigen::MatrixXd X;
        const bool ok = solver.solve(X);
        double e = 0; int ref_ok = 0;
        if (!ok) {
            int bad = 0, first = 0;
            for (int s : solver.getStatus()) { if (s != 0 && !bad) first = s; bad += (s != 0); }
            fprintf(stderr, "BatchednHQP::solve failed (%s): %s; %d bad status words, first 0x%x\n", c.name, solver.getLastError().c_str(), bad, first);
            snprintf(buf, sizeof(buf), "\"%s_err\": -1, \"%s_ok\": 0, \"%s_ref_ok\": 0, ", c.name, c.name, c.name); js += buf;
            continue;
        }
        if (!strcmp(c.name, "default")) {
            Xdef = X;
            // host share of a solve: the per-level preparation (SVD, damping, H, g, projected rows) against the C-ABI calls
            snprintf(buf, sizeof(buf), "\"prepare_us_per_stack\": %.2f, \"engine_us_per_stack\": %.2f, ",
                     1e6 * solver.getLastPrepareSeconds() / count, 1e6 * solver.getLastEngineSeconds() / count); js += buf;
        }
        else { snprintf(buf, sizeof(buf), "\"%s_moved\": %.3e, ", c.name, (X - Xdef).cwiseAbs().maxCoeff()); js += buf; }
        for (int i = 0; i < count; ++i) {
            if (ref[(size_t)i * (1 + n)] < 0.5) continue;
            ++ref_ok;
            for (int j = 0; j < n; ++j) e = std::max(e, std::fabs(X(j, i) - ref[(size_t)i * (1 + n) + 1 + j]));
        }
        snprintf(buf, sizeof(buf), "\"%s_err\": %.3e, \"%s_ok\": %d, \"%s_ref_ok\": %d, ", c.name, e, c.name, ok ? 1 : 0, c.name, ref_ok);
        js += buf;
    }
    // the reference's own signature, a batch of one, through the Solver<> base
    {
        BN::Base::SolverPtr one = std::make_shared<BN>(stacks[0], bounds[0], 2e2, solvers::solver_back_ends::qpOASES);
        nhqp_solve_blobs(&d, &L, blobs.data(), 1, 0.05, 1, 1, ref.data());
        Eigen::VectorXd x(n);
        const bool ok = one->solve(x);
        double e = 0;
        for (int j = 0; j < n; ++j) e = std::max(e, std::fabs(x[j] - ref[1 + j]));
        snprintf(buf, sizeof(buf), "\"single_err\": %.3e, \"single_ok\": %d, ", e, ok ? 1 : 0);
        js += buf;
    }
    snprintf(buf, sizeof(buf), "\"layers\": [%d, %d, %d], \"shape\": \"%s\", \"count\": %d}", solver.getFreeVariables(0),
             solver.getNumberOfTasks() > 1 ? solver.getFreeVariables(1) : 0, solver.getNumberOfTasks() > 2 ? solver.getFreeVariables(2) : 0, argv[2], count);
    js += buf;
    printf("\n%s\n", js.c_str());
    return 0;
}

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void on_crash(int sig)
{
    void* frames[48];
    const int n = backtrace(frames, 48);
    const char msg[] = "test_dropin: fatal signal, backtrace:\n";
    write(2, msg, sizeof(msg) - 1);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char** argv)
{
    signal(SIGSEGV, on_crash); signal(SIGABRT, on_crash); signal(SIGFPE, on_crash);
    if (argc >= 2 && !strcmp(argv[1], "surface")) return surface(argc, argv);
    if (argc >= 2 && !strcmp(argv[1], "nhqp")) return nhqp(argc, argv);
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
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 2\n");
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
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 2b\n");
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
    if (getenv("OSB_TEST_TRACE")) fprintf(stderr, "scenario 3\n");
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

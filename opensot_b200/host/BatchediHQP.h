/*
 * BatchediHQP.h -- batched front-end: B structurally identical OpenSoT stacks, one GPU solve.
 *
 * The batched counterpart of OpenSoT::solvers::iHQP (include/OpenSoT/solvers/iHQP.h:44-264) and,
 * like it, an OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> (include/OpenSoT/Solver.h:29-116).
 * It takes the SAME objects an existing program already has -- OpenSoT::AutoStack::Ptr built with the
 * Math-of-Tasks operators (+ * % / <<, include/OpenSoT/utils/AutoStack.h:115-336), or a Stack with
 * bounds / global constraints -- one per robot instance / MPC sample, and replaces
 *     for each instance: iHQP::solve()         src/solvers/iHQP.cpp:263-358
 * by ONE call into the C ABI (include/opensot_b200.h). Every iHQP constructor (iHQP.h:59-108) exists
 * twice: with the reference's exact signature (batch of one: existing code compiles unchanged with the
 * class name swapped) and with a std::vector of stacks as first argument (the batch).
 *
 * Per cycle the user still calls AutoStack::update() on every stack (that is where the robot model
 * produces A and b; updateAndSolve() does it, optionally on several threads). What this class
 * replaces on the host is the piling: Aggregated::generateAll / MatrixPiler copies
 * (src/tasks/Aggregated.cpp:113-132, src/constraints/Aggregated.cpp:73-257,
 * include/OpenSoT/utils/Piler.h:102-163) become writes of every task's rows at fixed offsets of the
 * instance blob (threads over instances, page-locked memory), and the BackEnd::updateTask /
 * updateConstraints / updateBounds deep copies (src/solvers/BackEnd.cpp:19-104) disappear.
 *
 * Header-only on purpose: it is compiled against whatever OpenSoT headers the host program uses, so
 * the Task / Constraint / AutoStack / boost::any ABI is the program's own.
 * Error behaviour mirrors the reference: constructors throw std::runtime_error (iHQP.cpp:24),
 * solve() returns false if any instance failed (iHQP.cpp:279,335-347) and the per-instance
 * osb_status words stay available through getStatus().
 */
#ifndef OPENSOT_B200_BATCHED_IHQP_H_
#define OPENSOT_B200_BATCHED_IHQP_H_

#include <OpenSoT/Constraint.h>
#include <OpenSoT/Solver.h>
#include <OpenSoT/Task.h>
#include <OpenSoT/solvers/BackEnd.h>
#include <OpenSoT/solvers/BackEndFactory.h>
#include <OpenSoT/utils/AutoStack.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <limits>
#include <list>
#include <mutex>
#include <stdexcept>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "opensot_b200.h"

namespace OpenSoT { namespace b200 {

/* a few persistent worker threads: a control cycle cannot afford thread creation (tens of
 * microseconds each) in front of a 0.2 ms GPU call */
class WorkerPool {
public:
    explicit WorkerPool(int threads) : _stop(false), _gen(0), _pending(0)
    {
        for (int t = 0; t < threads; ++t) _workers.emplace_back([this, t] { loop(t); });
    }
    ~WorkerPool()
    {
        { std::lock_guard<std::mutex> g(_m); _stop = true; ++_gen; }
        _cv.notify_all();
        for (auto& w : _workers) w.join();
    }
    int size() const { return (int)_workers.size(); }
    /* fn(i) for i in [0, count), dynamically chunked; the caller works too; returns when all are done */
    void run(int count, const std::function<void(int)>& fn)
    {
        if (_workers.empty() || count < 2 * 16) { for (int i = 0; i < count; ++i) fn(i); return; }
        _fn = &fn; _count = count; _next.store(0);
        { std::lock_guard<std::mutex> g(_m); _pending = (int)_workers.size(); ++_gen; }
        _cv.notify_all();
        work();
        std::unique_lock<std::mutex> l(_m);
        _done.wait(l, [this] { return _pending == 0; });
    }

private:
    void work()
    {
        for (;;) {
            const int lo = _next.fetch_add(16);
            if (lo >= _count) return;
            const int hi = lo + 16 < _count ? lo + 16 : _count;
            for (int i = lo; i < hi; ++i) (*_fn)(i);
        }
    }
    void loop(int)
    {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> l(_m);
                _cv.wait(l, [&] { return _gen != seen; });
                seen = _gen;
                if (_stop) return;
            }
            work();
            { std::lock_guard<std::mutex> g(_m); if (--_pending == 0) _done.notify_one(); }
        }
    }
    std::vector<std::thread> _workers;
    std::mutex _m;
    std::condition_variable _cv, _done;
    bool _stop;
    unsigned long _gen;
    int _pending, _count = 0;
    std::atomic<int> _next{0};
    const std::function<void(int)>* _fn = nullptr;
};

class BatchediHQP : public OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> {
public:
    typedef std::shared_ptr<BatchediHQP> Ptr;
    typedef OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> Base;
    typedef Base::TaskType TaskType;
    typedef Base::ConstraintType ConstraintType;
    typedef Base::TaskPtr TaskPtr;
    typedef Base::ConstraintPtr ConstraintPtr;
    typedef Base::Stack Stack;
    typedef OpenSoT::solvers::solver_back_ends solver_back_ends;
    typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> RowMat;

    /* ---- batched constructors: one entry per robot instance / MPC sample -------------------------
     * eps_regularisation is the same FACTOR iHQP takes (default 2e2, iHQP.h:32) */
    BatchediHQP(const std::vector<OpenSoT::AutoStack::Ptr>& stacks, const double eps_regularisation = 2e2,
                const int device = 0)
        : Base(first_stack(stacks), stacks.empty() ? ConstraintPtr() : stacks[0]->getBounds()), _eps_factor(eps_regularisation)
    {
        for (auto& s : stacks) {
            Instance in;
            in.autostack = s; in.tasks = s->getStack(); in.reg = s->getRegularisationTask();
            _inst.push_back(in);
        }
        init(device, "");
    }
    BatchediHQP(std::vector<Stack>& stacks, const double eps_regularisation = 2e2, const int device = 0)
        : Base(first_of(stacks)), _eps_factor(eps_regularisation)
    {
        for (auto& s : stacks) { Instance in; in.tasks = s; _inst.push_back(in); }
        init(device, "");
    }
    BatchediHQP(std::vector<Stack>& stacks, const std::vector<ConstraintPtr>& bounds,
                const double eps_regularisation = 2e2, const int device = 0)
        : Base(first_of(stacks), bounds.empty() ? ConstraintPtr() : bounds[0]), _eps_factor(eps_regularisation)
    {
        if (bounds.size() != stacks.size()) throw std::runtime_error("BatchediHQP: one bounds object per stack is needed");
        for (size_t i = 0; i < stacks.size(); ++i) { Instance in; in.tasks = stacks[i]; in.bounds = bounds[i]; _inst.push_back(in); }
        init(device, " with bounds");
    }
    BatchediHQP(std::vector<Stack>& stacks, const std::vector<ConstraintPtr>& bounds,
                const std::vector<ConstraintPtr>& globalConstraints, const double eps_regularisation = 2e2, const int device = 0)
        : Base(first_of(stacks), bounds.empty() ? ConstraintPtr() : bounds[0],
               globalConstraints.empty() ? ConstraintPtr() : globalConstraints[0]), _eps_factor(eps_regularisation)
    {
        if (bounds.size() != stacks.size() || globalConstraints.size() != stacks.size())
            throw std::runtime_error("BatchediHQP: one bounds and one globalConstraints object per stack are needed");
        for (size_t i = 0; i < stacks.size(); ++i) {
            Instance in; in.tasks = stacks[i]; in.bounds = bounds[i]; in.global = globalConstraints[i];
            _inst.push_back(in);
        }
        init(device, " with bounds");
    }

    /* ---- the reference's own constructor signatures (iHQP.h:59-108): a batch of one. The back-end
     * arguments are accepted for source compatibility; every level is solved by the B200 engine ---- */
    BatchediHQP(OpenSoT::AutoStack& stack_of_tasks, const double eps_regularisation = 2e2,
                const solver_back_ends = solver_back_ends::qpOASES)
        : Base(stack_of_tasks.getStack(), stack_of_tasks.getBounds()), _eps_factor(eps_regularisation)
    {
        Instance in;
        in.autostack_raw = &stack_of_tasks; in.tasks = stack_of_tasks.getStack(); in.reg = stack_of_tasks.getRegularisationTask();
        _inst.push_back(in);
        init(0, "");
    }
    BatchediHQP(OpenSoT::AutoStack& stack_of_tasks, const double eps_regularisation, const std::vector<solver_back_ends>)
        : BatchediHQP(stack_of_tasks, eps_regularisation, solver_back_ends::qpOASES) {}
    BatchediHQP(Stack& stack_of_tasks, const double eps_regularisation = 2e2, const solver_back_ends = solver_back_ends::qpOASES)
        : Base(stack_of_tasks), _eps_factor(eps_regularisation)
    {
        Instance in; in.tasks = stack_of_tasks; _inst.push_back(in);
        init(0, "");
    }
    BatchediHQP(Stack& stack_of_tasks, const double eps_regularisation, const std::vector<solver_back_ends>)
        : BatchediHQP(stack_of_tasks, eps_regularisation, solver_back_ends::qpOASES) {}
    BatchediHQP(Stack& stack_of_tasks, ConstraintPtr bounds, const double eps_regularisation = 2e2,
                const solver_back_ends = solver_back_ends::qpOASES)
        : Base(stack_of_tasks, bounds), _eps_factor(eps_regularisation)
    {
        Instance in; in.tasks = stack_of_tasks; in.bounds = bounds; _inst.push_back(in);
        init(0, " with bounds");
    }
    BatchediHQP(Stack& stack_of_tasks, ConstraintPtr bounds, const double eps_regularisation, const std::vector<solver_back_ends>)
        : BatchediHQP(stack_of_tasks, bounds, eps_regularisation, solver_back_ends::qpOASES) {}
    BatchediHQP(Stack& stack_of_tasks, ConstraintPtr bounds, ConstraintPtr globalConstraints,
                const double eps_regularisation = 2e2, const solver_back_ends = solver_back_ends::qpOASES)
        : Base(stack_of_tasks, bounds, globalConstraints), _eps_factor(eps_regularisation)
    {
        Instance in; in.tasks = stack_of_tasks; in.bounds = bounds; in.global = globalConstraints; _inst.push_back(in);
        init(0, " with bounds");
    }
    BatchediHQP(Stack& stack_of_tasks, ConstraintPtr bounds, ConstraintPtr globalConstraints, const double eps_regularisation,
                const std::vector<solver_back_ends>)
        : BatchediHQP(stack_of_tasks, bounds, globalConstraints, eps_regularisation, solver_back_ends::qpOASES) {}

    ~BatchediHQP()
    {
        osb_engine_destroy(_engine);
        osb_host_free(_blobs); osb_host_free(_x); osb_host_free(_xlev); osb_host_free(_status_pl);
    }
    BatchediHQP(const BatchediHQP&) = delete;
    BatchediHQP& operator=(const BatchediHQP&) = delete;

    /* ---- solving --------------------------------------------------------------------------------
     * solutions: n x B, column i = joint velocities of instance i (iHQP::solve for every instance) */
    bool solve(Eigen::MatrixXd& solutions)
    {
        const int B = (int)_inst.size(), n = _desc.n;
        if (!pack_all()) return false;
        if (osb_solve_host(_engine, _blobs, B, _x, _status_pl, _store_levels ? _xlev : nullptr) != OSB_SUCCESS) {
            _error = osb_last_error();
            return false;
        }
        std::memcpy(_status.data(), _status_pl, sizeof(int32_t) * (size_t)B);
        _solved_once = true;
        if (solutions.rows() != n || solutions.cols() != B) solutions.resize(n, B);
        /* an n x B column-major Eigen matrix is B consecutive solution vectors */
        std::memcpy(solutions.data(), _x, sizeof(double) * (size_t)n * B);
        for (int i = 0; i < B; ++i)
            if (_status[i] != OSB_OK) return false;
        return true;
    }
    /* Solver<>::solve (Solver.h:84; iHQP.cpp:263): the solution vector of a batch of one; for a larger
     * batch the B solutions one after another (length n B) */
    bool solve(Eigen::VectorXd& solution) override
    {
        const bool ok = solve(_X);
        if (solution.size() != _X.size()) solution.resize(_X.size());
        std::memcpy(solution.data(), _X.data(), sizeof(double) * (size_t)_X.size());
        return ok;
    }
    /* AutoStack::update() on every instance (src/utils/AutoStack.cpp:385-393), then solve */
    bool updateAndSolve(Eigen::MatrixXd& solutions)
    {
        update();
        return solve(solutions);
    }
    /* What AutoStack::update() does, for every instance: bounds aggregate, every level, regularisation
     * task. Serial by default: the reference library is not re-entrant (static ids in
     * tasks/Aggregated.h:113, constraints/Aggregated.h:99, BilateralConstraint.h:35);
     * setUpdateThreads(k > 1) runs instances on k threads for stacks whose tasks are thread-safe. */
    void update()
    {
        auto one = [this](int i) {
            Instance& in = _inst[i];
            if (in.autostack) in.autostack->update();
            else if (in.autostack_raw) in.autostack_raw->update();
            else {
                if (in.bounds) in.bounds->update();
                if (in.global) in.global->update();
                for (auto& t : in.tasks) t->update();
            }
        };
        if (_update_threads > 1) pool().run((int)_inst.size(), one);
        else for (int i = 0; i < (int)_inst.size(); ++i) one(i);
    }
    void setUpdateThreads(const int k) { _update_threads = k; }
    /* threads that pile instances into blobs (default: hardware threads, at most 16; 1 = caller only) */
    void setPackThreads(const int k) { _pack_threads = k < 1 ? 1 : k; _pool.reset(); }

    /* ---- the rest of iHQP's surface (iHQP.h:118-193) ---------------------------------------------- */
    unsigned int getNumberOfTasks() { return (unsigned int)_desc.n_levels; }
    unsigned int getBatchSize() const { return (unsigned int)_inst.size(); }
    void setActiveStack(const unsigned int i, const bool flag)
    {
        if (i < _active.size()) { _active[i] = flag ? 1 : 0; push_active(); }
    }
    void activateAllStacks() { _active.assign(_active.size(), 1); push_active(); }
    std::string getBackEndName(const unsigned int i)
    {
        if (i >= (unsigned)_desc.n_levels) return "";        /* iHQP.cpp:417-425 */
        return "B200";
    }
    /* options of level i as a boost::any holding an osb_options (the engine keeps ONE set of options for
     * the whole cascade: setting them through any level sets them for all); false if i is out of range or
     * the any holds something else (iHQP.cpp:360-378) */
    bool setOptions(const unsigned int i, const boost::any& opt)
    {
        if (i >= (unsigned)_desc.n_levels) return false;
        try {
            osb_options o = boost::any_cast<osb_options>(opt);
            o.store_levels = _store_levels ? 1 : 0;
            return osb_set_options(_engine, &o) == OSB_SUCCESS;
        } catch (...) { _error = "setOptions: the boost::any must hold an osb_options"; return false; }
    }
    bool getOptions(const unsigned int i, boost::any& opt)
    {
        if (i >= (unsigned)_desc.n_levels) return false;
        osb_options o;
        if (osb_get_options(_engine, &o) != OSB_SUCCESS) return false;
        opt = o;
        return true;
    }
    /* value of level i's objective at its optimum, 1/2 x'Hx + g'x with H = A'WA (+ Hr) + eps I,
     * g = -A'Wb + c (+ gr): what BackEnd::getObjective returns for qpOASES (QPOasesBackEnd.cpp:334-337),
     * for instance `instance` (iHQP.cpp:380-388). Needs the level solutions: see setStoreLevels. */
    bool getObjective(const unsigned int i, double& val, const unsigned int instance = 0)
    {
        if (i >= (unsigned)_desc.n_levels || instance >= _inst.size() || !_store_levels || !_solved_once) return false;
        const int n = _desc.n;
        Eigen::Map<const Eigen::VectorXd> x(_xlev + ((size_t)instance * _desc.n_levels + i) * n, n);
        Instance& in = _inst[instance];
        auto cost = [&](TaskType& t) {
            const Eigen::VectorXd Ax = t.getA() * x;
            const Eigen::VectorXd WAx = t.getWeight() * Ax;
            double v = 0.5 * Ax.dot(WAx) - WAx.dot(t.getb());
            if (t.getc().size() == n) v += t.getc().dot(x);
            return v;
        };
        val = cost(*in.tasks[i]) + 0.5 * osb_get_eps_level(_engine, (int)i) * x.squaredNorm();
        if (in.reg) val += cost(*in.reg);
        return true;
    }
    /* OVERWRITES the absolute diagonal shift of level i / of every level (iHQP.cpp:437-459,
     * QPOasesBackEnd.cpp:358-372: the value is absolute, not the constructor's factor) */
    bool setEpsRegularisation(const double eps, const unsigned int i)
    {
        if (i >= (unsigned)_desc.n_levels || eps < 0.0) return false;
        return osb_set_eps_level(_engine, (int)i, eps) == OSB_SUCCESS;
    }
    bool setEpsRegularisation(const double eps)
    {
        if (eps < 0.0) return false;
        return osb_set_eps_level(_engine, -1, eps) == OSB_SUCCESS;
    }
    double getEpsRegularisation(const unsigned int i) { return osb_get_eps_level(_engine, (int)i); }

    /* A read-only BackEnd view of level i (iHQP.cpp:427-436): getSolution(), getH(), getg(), getA(),
     * getlA(), getuA(), getl(), getu(), getObjective(), getEpsRegularisation() report what the reference's
     * back-end of that level would hold for instance `instance` after the last solve() (H, g as
     * iHQP::computeCostFunction forms them, A = [rows of the constraints ; optimality rows A_j with
     * lA = uA = A_j x_j*], iHQP.cpp:285-332). It cannot be solved on its own. */
    class LevelView : public OpenSoT::solvers::BackEnd {
    public:
        LevelView(int n, int nc) : BackEnd(n, nc), _obj(0.0), _eps(0.0) {}
        bool initProblem(const Eigen::MatrixXd&, const Eigen::VectorXd&, const Eigen::MatrixXd&, const Eigen::VectorXd&,
                         const Eigen::VectorXd&, const Eigen::VectorXd&, const Eigen::VectorXd&) override { return false; }
        bool solve() override { return false; }
        boost::any getOptions() override { return _opt; }
        void setOptions(const boost::any&) override {}
        double getObjective() override { return _obj; }
        double getEpsRegularisation() override { return _eps; }
        friend class BatchediHQP;
    private:
        double _obj, _eps;
        osb_options _opt;
    };
    bool getBackEnd(const unsigned int i, OpenSoT::solvers::BackEnd::Ptr& back_end, const unsigned int instance = 0)
    {
        if (i >= (unsigned)_desc.n_levels || instance >= _inst.size() || !_store_levels || !_solved_once) return false;
        const int n = _desc.n, L = _desc.n_levels;
        Instance& in = _inst[instance];
        const double* blob = _blobs + (size_t)instance * _layout.stride;
        /* rows of this level's QP: task-local rows, global rows, optimality rows of the earlier levels */
        const std::list<ConstraintPtr> glist = global_list(in);
        const int grows = total_rows(measure(glist), n, false);                          /* rows this instance has NOW: the */
        const int lrows = total_rows(measure(in.tasks[i]->getConstraints()), n, true);   /* unused slots are not reported    */
        int rows = grows + lrows;
        for (unsigned j = 0; j < i; ++j) rows += _desc.m[j];
        std::shared_ptr<LevelView> v = std::make_shared<LevelView>(n, rows);
        TaskType& t = *in.tasks[i];
        v->_H = t.getA().transpose() * t.getWeight() * t.getA();
        v->_g = -t.getA().transpose() * t.getWeight() * t.getb();
        if (t.getc().size() == n) v->_g += t.getc();
        if (in.reg) {
            TaskType& r = *in.reg;
            v->_H += r.getA().transpose() * r.getWeight() * r.getA();
            v->_g += -r.getA().transpose() * r.getWeight() * r.getb();
        }
        v->_A.resize(rows, n); v->_lA.resize(rows); v->_uA.resize(rows);
        int at = 0;
        if (grows + lrows > 0) {
            std::vector<double> C((size_t)(grows + lrows) * n), lo(grows + lrows), hi(grows + lrows);
            /* the level's own constraints first, then the global ones (iHQP.cpp:189-197) */
            gather_rows(in.tasks[i]->getConstraints(), n, true, nullptr, C.data(), lo.data(), hi.data());
            gather_rows(glist, n, false, nullptr, C.data() + (size_t)lrows * n, lo.data() + lrows, hi.data() + lrows);
            for (int r = 0; r < grows + lrows; ++r, ++at) {
                v->_A.row(at) = Eigen::Map<const Eigen::VectorXd>(C.data() + (size_t)r * n, n).transpose();
                v->_lA[at] = lo[r]; v->_uA[at] = hi[r];
            }
        }
        for (unsigned j = 0; j < i; ++j) {
            Eigen::Map<const Eigen::VectorXd> xj(_xlev + ((size_t)instance * L + j) * n, n);
            const Eigen::MatrixXd& Aj = in.tasks[j]->getA();
            const Eigen::VectorXd rhs = Aj * xj;
            for (int r = 0; r < Aj.rows(); ++r, ++at) {
                if (_active[j]) { v->_A.row(at) = Aj.row(r); v->_lA[at] = v->_uA[at] = rhs[r]; }
                else { v->_A.row(at).setZero(); v->_lA[at] = -1.0; v->_uA[at] = 1.0; }   /* iHQP.cpp:303-309 */
            }
        }
        if (_n_bound_cons > 0) {
            v->_l = Eigen::Map<const Eigen::VectorXd>(blob + _layout.l[0], n);
            v->_u = Eigen::Map<const Eigen::VectorXd>(blob + _layout.u[0], n);
        }
        v->_solution = Eigen::Map<const Eigen::VectorXd>(_xlev + ((size_t)instance * L + i) * n, n);
        v->_eps = osb_get_eps_level(_engine, (int)i);
        getObjective(i, v->_obj, instance);
        osb_get_options(_engine, &v->_opt);
        back_end = v;
        return true;
    }
    /* Solver<>::log -> _log (Solver.h:46-115), what iHQP::_log writes (iHQP.cpp:403-415): Hr, gr of a regularisation
     * task, then BackEnd::log of every level (BackEnd.cpp:111-130: H_i, g_i, A_i, lA_i, uA_i, l_i, u_i, solution_i under
     * the solver-id prefix) -- for ONE instance of the batch (setLogInstance, default 0), through the level views of
     * getBackEnd. The batch as a whole is logged by opensot_b200/qplog.py. Needs level storage and a solve. */
    void setLogInstance(const unsigned int instance) { _log_instance = instance; }
    void _log(XBot::MatLogger2::Ptr logger, const std::string& prefix) override
    {
        if (!logger || _log_instance >= _inst.size()) return;
        Instance& in = _inst[_log_instance];
        if (in.reg) {
            TaskType& r = *in.reg;
            const Eigen::MatrixXd Hr = r.getA().transpose() * r.getWeight() * r.getA();
            Eigen::VectorXd gr = -r.getA().transpose() * r.getWeight() * r.getb();
            if (r.getc().size() == gr.size()) gr += r.getc();
            if (Hr.rows() > 0) logger->add("Hr", Hr);
            if (gr.size() > 0) logger->add("gr", gr);
        }
        for (unsigned i = 0; i < (unsigned)_desc.n_levels; ++i) {
            OpenSoT::solvers::BackEnd::Ptr be;
            if (getBackEnd(i, be, _log_instance)) be->log(logger, (int)i, prefix);
        }
    }
    /* per-level solutions travel back with every solve (3 x n doubles per instance for a 3-level stack
     * against ~600 uploaded); switch off when neither getObjective nor getBackEnd is used */
    void setStoreLevels(const bool on)
    {
        _store_levels = on;
        osb_options o;
        if (osb_get_options(_engine, &o) == OSB_SUCCESS) { o.store_levels = on ? 1 : 0; osb_set_options(_engine, &o); }
    }
    /* solution of level `level`: n x B (after a solve with level storage on) */
    bool getLevelSolutions(const unsigned int level, Eigen::MatrixXd& solutions) const
    {
        if (level >= (unsigned)_desc.n_levels || !_store_levels || !_solved_once) return false;
        const int B = (int)_inst.size(), n = _desc.n, L = _desc.n_levels;
        solutions.resize(n, B);
        for (int i = 0; i < B; ++i)
            solutions.col(i) = Eigen::Map<const Eigen::VectorXd>(_xlev + ((size_t)i * L + level) * n, n);
        return true;
    }

    /* spare inequality-row slots a constraint gets when it outgrows the ones it has (a ConvexHull that gains an edge,
     * ConvexHull.cpp:47-52): every growth rebuilds the engine, headroom makes that rare. Default 0: stacks whose row
     * counts never change carry no padding */
    void setRowHeadroom(const int rows) { _row_headroom = rows < 0 ? 0 : rows; }
    /* engine rebuilds since construction (structure drift: rows outgrowing their slots, equalities that stopped being
     * ones, identity tasks that stopped being ones) */
    int getRebuilds() const { return _rebuilds; }
    const std::vector<int32_t>& getStatus() const { return _status; }
    const std::string& getLastError() const { return _error; }
    const osb_stack_desc& getDescriptor() const { return _desc; }
    const char* getKernelName() const { return osb_engine_kernel_name(_engine); }
    /* times of the last solve(): piling on the host, and the C-ABI call (H2D + kernel + D2H) */
    double getLastPackSeconds() const { return _t_pack; }

private:
    struct Instance {
        OpenSoT::AutoStack::Ptr autostack;
        OpenSoT::AutoStack* autostack_raw = nullptr;
        Stack tasks;
        ConstraintPtr bounds, global;
        TaskPtr reg;
    };
    static Stack& first_stack(const std::vector<OpenSoT::AutoStack::Ptr>& stacks)
    {
        if (stacks.empty() || !stacks[0]) throw std::runtime_error("BatchediHQP: no stacks given");
        return stacks[0]->getStack();
    }
    static Stack& first_of(std::vector<Stack>& stacks)
    {
        if (stacks.empty()) throw std::runtime_error("BatchediHQP: no stacks given");
        return stacks[0];
    }
    static bool is_zero(const Eigen::VectorXd& v) { return v.size() == 0 || v.cwiseAbs().maxCoeff() == 0.0; }

    /* The constraints every level of an instance sees, as the reference assembles them:
     *  AutoStack : everything `stack << ...` collected (AutoStack.cpp:303-316; iHQP receives it as one
     *              constraints::Aggregated, iHQP.cpp:12-14, whose members all contribute rows AND bounds);
     *  Stack     : globalConstraints if given, else `bounds` when it carries rows; `bounds` again for its
     *              box when it is a pure bound (iHQP.cpp:187-195, 230-234). */
    std::list<ConstraintPtr> global_list(Instance& in) const
    {
        if (in.autostack) return in.autostack->getBoundsList();
        if (in.autostack_raw) return in.autostack_raw->getBoundsList();
        std::list<ConstraintPtr> l;
        if (in.global) l.push_back(in.global);
        else if (in.bounds && in.bounds->isConstraint()) l.push_back(in.bounds);
        if (in.bounds && in.bounds->isBound()) l.push_back(in.bounds);
        return l;
    }

    void init(const int device, const char* what)
    {
        if (_inst.empty()) throw std::runtime_error("BatchediHQP: no stacks given");
        if (!prepare(device)) throw std::runtime_error(std::string("Can Not initizalize SoT") + what + "! " + _error);
    }

    bool prepare(const int device)
    {
        Instance& i0 = _inst[0];
        if (i0.tasks.empty() || i0.tasks.size() > OSB_MAX_LEVELS) { _error = "unsupported number of levels"; return false; }
        std::memset(&_desc, 0, sizeof(_desc));
        const int n = (int)i0.tasks[0]->getXSize();
        _desc.n = n;
        _desc.n_levels = (int)i0.tasks.size();
        _desc.eps = _eps_factor * OSB_QPOASES_BASE_EPS;
        for (size_t k = 0; k < i0.tasks.size(); ++k) {
            TaskPtr t = i0.tasks[k];
            _desc.m[k] = (int)t->getA().rows();
            /* A == I on every instance (Postural-like, tasks/velocity/Postural.cpp): the engine's identity level, no
             * matrix stored or factored; should A stop being the identity the engine is rebuilt (pack_all) */
            bool a_ident = _desc.m[k] == n && !(_not_identity.size() > k && _not_identity[k]);
            bool diag = true, ident = true, has_c = false;
            for (auto& in : _inst) {
                if (in.tasks.size() != i0.tasks.size()) { _error = "stacks differ in shape"; return false; }
                TaskPtr tk = in.tasks[k];
                if ((int)tk->getA().rows() != _desc.m[k] || (int)tk->getXSize() != n) { _error = "stacks differ in shape"; return false; }
                const Eigen::MatrixXd& W = tk->getWeight();
                if (!is_diagonal(W)) diag = false;
                if (!W.isIdentity(0.0)) ident = false;
                if (!is_zero(tk->getc())) has_c = true;
                if (a_ident && !tk->getA().isIdentity(0.0)) a_ident = false;
            }
            /* what a level has been once it stays able to be (a weight that gained off-diagonal terms, a linear cost
             * term that appeared after construction): pack() reports the drift, pack_all() rebuilds */
            if (_relax.size() != i0.tasks.size()) _relax.assign(i0.tasks.size(), Relax());
            if (_relax[k].weight >= 1) ident = false;
            if (_relax[k].weight >= 2) diag = false;
            if (_relax[k].c) has_c = true;
            _desc.task_kind[k] = a_ident ? OSB_TASK_IDENTITY : OSB_TASK_DENSE;
            _desc.weight_kind[k] = ident ? OSB_W_IDENTITY : (diag ? OSB_W_DIAG : OSB_W_DENSE);
            _desc.has_c[k] = has_c ? 1 : 0;
            /* task-local constraints: every member of the list owns as many inequality-row slots as the most rows it
             * has on any instance (now or earlier), see Seg */
            if (_lcap.size() != i0.tasks.size()) _lcap.assign(i0.tasks.size(), std::vector<Seg>());
            for (auto& in : _inst)
                if (!merge_caps(_lcap[k], measure(in.tasks[k]->getConstraints()))) { _error = "stacks differ in task-local constraints"; return false; }
            _desc.c_loc[k] = total_rows(_lcap[k], n, true);
        }
        /* global constraints and bounds. Every member contributes its rows AND its box, whether or not it
         * is a pure bound (src/constraints/Aggregated.cpp:118-148 intersects l/u of every member) */
        _n_bound_cons = 0;
        for (auto& c : global_list(i0))
            if (c->hasBounds()) ++_n_bound_cons;
        for (auto& in : _inst)
            if (!merge_caps(_gcap, measure(global_list(in)))) { _error = "stacks differ in constraint rows"; return false; }
        const int rows = total_rows(_gcap, n, false);
        _desc.n_bound_sets = _n_bound_cons > 0 ? 1 : 0; /* intersected on the host while packing */
        /* equality rows are structural for the engine: a global row is eliminated as an equality if
         * bL == bU on every instance now; should that stop being true, the engine is rebuilt (pack_all) */
        _global_rows = rows;
        _row_is_eq.assign((size_t)rows, 1);
        std::vector<double> C((size_t)(rows > 0 ? rows : 1) * n), lo((size_t)rows + 1), hi((size_t)rows + 1);
        for (auto& in : _inst) {
            if (gather_rows(global_list(in), n, false, &_gcap, C.data(), lo.data(), hi.data()) != rows) { _error = "stacks differ in constraint rows"; return false; }
            for (int r = 0; r < rows; ++r)
                if (!(lo[r] == hi[r])) _row_is_eq[r] = 0;
        }
        /* (member, row within the member's slots) of every global row: what survives a rebuild with more slots */
        _row_key.clear();
        for (size_t j = 0; j < _gcap.size(); ++j)
            for (int i = 0; i < _gcap[j].eq + _gcap[j].in; ++i) _row_key.push_back(std::make_pair((int)j, i));
        for (int r = 0; r < rows; ++r)
            if (_never_eq.count(_row_key[(size_t)r])) _row_is_eq[r] = 0;
        _desc.c_eq = _desc.c_in = 0;
        for (int r = 0; r < rows; ++r) (_row_is_eq[r] ? _desc.c_eq : _desc.c_in)++;
        TaskPtr reg = i0.reg;
        if (reg) {
            _desc.m_reg = (int)reg->getA().rows();
            if (!is_diagonal(reg->getWeight())) { _error = "a regularisation task needs a diagonal weight"; return false; }
            _desc.reg_weight_kind = reg->getWeight().isIdentity(0.0) ? OSB_W_IDENTITY : OSB_W_DIAG;
        }
        if (osb_blob_layout(&_desc, &_layout) != OSB_SUCCESS) { _error = "stack shape not supported by the engine"; return false; }
        const int B = (int)_inst.size();
        if (_engine) { osb_engine_destroy(_engine); _engine = nullptr; }
        if (osb_engine_create(&_engine, &_desc, B, device) != OSB_SUCCESS) { _error = osb_last_error(); return false; }
        _device = device;
        osb_options o;
        osb_get_options(_engine, &o);
        o.store_levels = _store_levels ? 1 : 0;
        osb_set_options(_engine, &o);
        osb_host_free(_blobs); osb_host_free(_x); osb_host_free(_xlev); osb_host_free(_status_pl);
        _blobs = (double*)osb_host_alloc((int64_t)sizeof(double) * B * _layout.stride);
        _x = (double*)osb_host_alloc((int64_t)sizeof(double) * B * n);
        _xlev = (double*)osb_host_alloc((int64_t)sizeof(double) * B * n * _desc.n_levels);
        /* status words too: a device-to-host copy into pageable memory blocks the calling thread until the chunk it belongs
         * to has been solved, which serialises the copy / solve / copy-out pipeline of osb_solve_host */
        _status_pl = (int32_t*)osb_host_alloc((int64_t)sizeof(int32_t) * B);
        if (!_blobs || !_x || !_xlev || !_status_pl) { _error = "page-locked host memory could not be allocated"; return false; }
        std::memset(_blobs, 0, sizeof(double) * (size_t)B * _layout.stride);
        _status.assign((size_t)B, 0);
        if (_active.size() != (size_t)_desc.n_levels) _active.assign((size_t)_desc.n_levels, 1);
        push_active();
        return true;
    }

    static bool is_diagonal(const Eigen::MatrixXd& W)
    {
        for (int j = 0; j < W.cols(); ++j)
            for (int i = 0; i < W.rows(); ++i)
                if (i != j && W(i, j) != 0.0) return false;
        return true;
    }

    /* Rows of one member of a constraint list. Equality rows and the box are structural; the number of inequality rows may
     * differ between instances and change at run time (constraints/velocity/ConvexHull.cpp:47-52: one row per hull edge) --
     * the reference then rebuilds its QP object (QPOasesBackEnd.cpp:229-244). The engine has fixed shapes: a member owns
     * `in` inequality-row SLOTS, the most rows it has had on any instance so far; unused slots hold the null row
     * -inf <= 0 x <= +inf (the convention CollisionAvoidance.cpp:104-107 already uses), which no iterate violates. A member
     * that outgrows its slots makes pack_all() rebuild the engine with more (plus setRowHeadroom). */
    struct Seg { int eq, in, box; };
    template <class List>
    static std::vector<Seg> measure(const List& list)
    {
        std::vector<Seg> v;
        for (auto& cp : list) v.push_back(Seg{(int)cp->getAeq().rows(), (int)cp->getAineq().rows(), cp->hasBounds() ? 1 : 0});
        return v;
    }
    /* slots := max(slots, rows now); false if the lists differ in members, equality rows or boxes */
    bool merge_caps(std::vector<Seg>& cap, const std::vector<Seg>& cur) const
    {
        const int headroom = _growing ? _row_headroom : 0;
        if (cap.empty() && !cur.empty()) { cap = cur; return true; }
        if (cap.size() != cur.size()) return false;
        for (size_t j = 0; j < cap.size(); ++j) {
            if (cap[j].eq != cur[j].eq || cap[j].box != cur[j].box) return false;
            if (cur[j].in > cap[j].in) cap[j].in = cur[j].in + headroom;
        }
        return true;
    }
    static int total_rows(const std::vector<Seg>& cap, const int n, const bool bounds_as_rows)
    {
        int rows = 0;
        for (auto& s : cap) rows += s.eq + s.in + ((bounds_as_rows && s.box) ? n : 0);
        return rows;
    }

    /* writes the rows of every constraint of the list in list order -- equality rows then inequality rows
     * of each member, what constraints::Aggregated::generateAll piles into Aineq / bLowerBound /
     * bUpperBound (src/constraints/Aggregated.cpp:150-237) -- and, when bounds_as_rows, n identity rows
     * for every member that has a box. With `caps` every member fills its slots (null rows behind its last
     * inequality row); returns the number of rows written, -1 if a member has more inequality rows than slots,
     * -2 if the list no longer has the members / equality rows / boxes it was built with. */
    template <class List>
    static int gather_rows(const List& list, const int n, const bool bounds_as_rows, const std::vector<Seg>* caps,
                           double* C, double* lo, double* hi)
    {
        const double inf = std::numeric_limits<double>::infinity();
        int r = 0;
        size_t j = 0;
        if (caps && caps->size() != (size_t)std::distance(list.begin(), list.end())) return -2;
        for (auto& cp : list) {
            ConstraintType& c = *cp;
            const Eigen::MatrixXd& Aeq = c.getAeq();
            const Eigen::MatrixXd& Ai = c.getAineq();
            if (caps) {
                const Seg& s = (*caps)[j];
                if ((int)Aeq.rows() != s.eq || (c.hasBounds() ? 1 : 0) != s.box) return -2;
                if ((int)Ai.rows() > s.in) return -1;
            }
            for (int i = 0; i < Aeq.rows(); ++i, ++r) { /* Aeq x = beq  ->  beq <= Aeq x <= beq */
                for (int k = 0; k < n; ++k) C[(size_t)r * n + k] = Aeq(i, k);
                lo[r] = hi[r] = c.getbeq()[i];
            }
            for (int i = 0; i < Ai.rows(); ++i, ++r) { /* one-sided rows get an infinite other side */
                for (int k = 0; k < n; ++k) C[(size_t)r * n + k] = Ai(i, k);
                lo[r] = c.getbLowerBound().size() ? c.getbLowerBound()[i] : -inf;
                hi[r] = c.getbUpperBound().size() ? c.getbUpperBound()[i] : inf;
            }
            if (caps)
                for (int i = (int)Ai.rows(); i < (*caps)[j].in; ++i, ++r) { /* unused slots: null rows */
                    for (int k = 0; k < n; ++k) C[(size_t)r * n + k] = 0.0;
                    lo[r] = -inf; hi[r] = inf;
                }
            if (bounds_as_rows && c.hasBounds()) {
                for (int i = 0; i < n; ++i, ++r) {
                    for (int k = 0; k < n; ++k) C[(size_t)r * n + k] = (i == k) ? 1.0 : 0.0;
                    lo[r] = c.getLowerBound().size() ? c.getLowerBound()[i] : -inf;
                    hi[r] = c.getUpperBound().size() ? c.getUpperBound()[i] : inf;
                }
            }
            ++j;
        }
        return r;
    }

    WorkerPool& pool()
    {
        if (!_pool) {
            int k = _pack_threads;
            if (k == 0) { k = (int)std::thread::hardware_concurrency(); if (k > 16) k = 16; if (k < 1) k = 1; }
            _pool.reset(new WorkerPool(k - 1));
        }
        return *_pool;
    }

    /* piling of every instance; a structural change that the engine can absorb by being rebuilt (an
     * equality row that stopped being one) triggers exactly that, anything else is an error */
    bool pack_all()
    {
        const auto t0 = std::chrono::steady_clock::now();
        const int B = (int)_inst.size();
        for (int attempt = 0; attempt < 4; ++attempt) {
            std::atomic<int> bad(0), eq_broken(0), ident_broken(0), grew(0), relaxed(0);
            std::vector<std::string> errs((size_t)pool().size() + 1);
            pool().run(B, [&](int i) {
                std::string err; int broke = -1;
                if (!pack(i, _blobs + (size_t)i * _layout.stride, err, broke)) {
                    if (broke == ROWS_GREW) grew.store(1);
                    else if (broke == RELAXED) relaxed.store(1);
                    else if (broke <= -2) {
                        ident_broken.store(1); std::lock_guard<std::mutex> g(_m_err);
                        if (_not_identity.size() < (size_t)_desc.n_levels) _not_identity.resize((size_t)_desc.n_levels, 0);
                        _not_identity[(size_t)(-broke - 2)] = 1;
                    } else if (broke >= 0) { eq_broken.store(1); std::lock_guard<std::mutex> g(_m_err); _broken_rows.push_back(broke); }
                    else { bad.store(1); std::lock_guard<std::mutex> g(_m_err); _error = err; }
                }
            });
            if (bad.load()) return false;
            if (!eq_broken.load() && !ident_broken.load() && !grew.load() && !relaxed.load()) { _t_pack = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); return true; }
            /* rows that were equalities at construction carry lo != hi now, an identity task matrix stopped
             * being one, a weight or a linear cost term left the class the engine was built for, or a constraint has more
             * inequality rows than slots: reclassify / re-measure and rebuild
             * the engine (what the reference does per back-end, QPOasesBackEnd.cpp:229-244) */
            for (int r : _broken_rows) _never_eq.insert(_row_key[(size_t)r]);
            _broken_rows.clear();
            _growing = grew.load() != 0;
            const bool ok = prepare(_device);
            _growing = false;
            if (!ok) return false;
            ++_rebuilds;
        }
        _error = "the structure of the stacks keeps changing (equality rows / identity tasks / constraint rows)";
        return false;
    }

    bool pack(const int inst, double* blob, std::string& err, int& broke)
    {
        Instance& s = _inst[inst];
        const int n = _desc.n;
        for (int k = 0; k < _desc.n_levels; ++k) {
            TaskType& t = *s.tasks[k];
            const int m = _desc.m[k];
            const Eigen::MatrixXd& A = t.getA();
            if (A.rows() != m || A.cols() != n) { err = "task changed size"; return false; }
            if (_desc.task_kind[k] == OSB_TASK_IDENTITY) {
                if (!A.isIdentity(0.0)) { broke = -2 - k; return false; }
            } else
                Eigen::Map<RowMat>(blob + _layout.A[k], m, n) = A;     /* column-major -> row-major block copy */
            std::memcpy(blob + _layout.b[k], t.getb().data(), sizeof(double) * m);
            const Eigen::MatrixXd& W = t.getWeight();
            if (_desc.weight_kind[k] == OSB_W_DIAG) {
                if (!is_diagonal(W)) { relax(k, 2, false); broke = RELAXED; return false; }
                for (int r = 0; r < m; ++r) blob[_layout.w[k] + r] = W(r, r);
            } else if (_desc.weight_kind[k] == OSB_W_DENSE)
                Eigen::Map<RowMat>(blob + _layout.w[k], m, m) = W;
            else if (!W.isIdentity(0.0)) { relax(k, is_diagonal(W) ? 1 : 2, false); broke = RELAXED; return false; }
            if (_desc.has_c[k]) {
                if (t.getc().size() == n) std::memcpy(blob + _layout.c[k], t.getc().data(), sizeof(double) * n);
                else std::memset(blob + _layout.c[k], 0, sizeof(double) * n);
            } else if (!is_zero(t.getc())) { relax(k, 0, true); broke = RELAXED; return false; }
            if (_desc.c_loc[k] > 0) {
                const int got = gather_rows(t.getConstraints(), n, true, &_lcap[(size_t)k], blob + _layout.Cl[k],
                                            blob + _layout.lCl[k], blob + _layout.uCl[k]);
                if (got == -1) { broke = ROWS_GREW; return false; }
                if (got != _desc.c_loc[k]) { err = "task-local constraints changed structure"; return false; }
            } else if (!t.getConstraints().empty() && total_rows(measure(t.getConstraints()), n, true) > 0) { broke = ROWS_GREW; return false; }
        }
        const std::list<ConstraintPtr> cons = global_list(s);
        /* bounds: l = max l_i, u = min u_i over every member that has a box (Aggregated.cpp:118-148) */
        if (_n_bound_cons > 0) {
            double* l = blob + _layout.l[0];
            double* u = blob + _layout.u[0];
            for (int i = 0; i < n; ++i) { l[i] = -std::numeric_limits<double>::infinity(); u[i] = -l[i]; }
            for (auto& c : cons) {
                if (!c->hasBounds()) continue;
                for (int i = 0; i < n; ++i) {
                    if (c->getLowerBound().size()) l[i] = std::max(l[i], c->getLowerBound()[i]);
                    if (c->getUpperBound().size()) u[i] = std::min(u[i], c->getUpperBound()[i]);
                }
            }
        }
        if (_global_rows > 0) {
            std::vector<double> tC((size_t)_global_rows * n), tlo((size_t)_global_rows), thi((size_t)_global_rows);
            const int got = gather_rows(cons, n, false, &_gcap, tC.data(), tlo.data(), thi.data());
            if (got == -1) { broke = ROWS_GREW; return false; }
            if (got != _global_rows) { err = "global constraints changed structure"; return false; }
            int e = 0, q = 0;
            for (int r = 0; r < _global_rows; ++r) {
                if (_row_is_eq[r]) {
                    if (!(tlo[r] == thi[r])) { broke = r; return false; }
                    std::memcpy(blob + _layout.Ceq + (size_t)e * n, tC.data() + (size_t)r * n, sizeof(double) * n);
                    blob[_layout.deq + e] = tlo[r];
                    ++e;
                } else {
                    std::memcpy(blob + _layout.Cin + (size_t)q * n, tC.data() + (size_t)r * n, sizeof(double) * n);
                    blob[_layout.lA + q] = tlo[r];
                    blob[_layout.uA + q] = thi[r];
                    ++q;
                }
            }
        }
        else if (total_rows(measure(cons), n, false) > 0) { broke = ROWS_GREW; return false; }
        if (_desc.m_reg > 0) {
            TaskPtr reg = s.reg;
            const int mr = _desc.m_reg;
            if (!reg || reg->getA().rows() != mr) { err = "regularisation task changed size"; return false; }
            if (!is_diagonal(reg->getWeight())) { err = "the regularisation task's weight stopped being diagonal"; return false; }
            Eigen::Map<RowMat>(blob + _layout.Ar, mr, n) = reg->getA();
            std::memcpy(blob + _layout.br, reg->getb().data(), sizeof(double) * mr);
            if (_desc.reg_weight_kind == OSB_W_DIAG)
                for (int r = 0; r < mr; ++r) blob[_layout.wr + r] = reg->getWeight()(r, r);
            else if (!reg->getWeight().isIdentity(0.0)) { err = "the regularisation task's weight stopped being the identity"; return false; }
        }
        return true;
    }

    void push_active() { if (_engine) osb_set_active_levels(_engine, _active.data()); }

    std::vector<Instance> _inst;
    osb_engine* _engine = nullptr;
    osb_stack_desc _desc;
    osb_layout _layout;
    double _eps_factor;
    int _device = 0;
    int _n_bound_cons = 0, _global_rows = 0;
    int _update_threads = 1, _pack_threads = 0;
    bool _store_levels = true, _solved_once = false;
    enum { ROWS_GREW = -1000, RELAXED = -1001 };   /* pack(): a constraint has more inequality rows than slots; a level needs a wider class */
    struct Relax { int weight = 0; bool c = false; };   /* weight: 1 = at least diagonal, 2 = dense */
    std::vector<Relax> _relax;
    void relax(const int k, const int weight, const bool c)
    {
        std::lock_guard<std::mutex> g(_m_err);
        if (_relax.size() <= (size_t)k) _relax.resize((size_t)k + 1);
        if (weight > _relax[(size_t)k].weight) _relax[(size_t)k].weight = weight;
        if (c) _relax[(size_t)k].c = true;
    }
    std::vector<char> _row_is_eq, _not_identity;
    std::vector<int> _broken_rows;
    std::vector<Seg> _gcap;                         /* row slots of the global constraint list, per member */
    std::vector<std::vector<Seg>> _lcap;            /* ... of every level's task-local list */
    std::vector<std::pair<int, int>> _row_key;      /* global row -> (member, row within its slots) */
    std::set<std::pair<int, int>> _never_eq;        /* rows that stopped being equalities once */
    int _row_headroom = 0, _rebuilds = 0;
    unsigned int _log_instance = 0;
    bool _growing = false;
    std::mutex _m_err;
    double *_blobs = nullptr, *_x = nullptr, *_xlev = nullptr;   /* page-locked */
    int32_t* _status_pl = nullptr;                               /* page-locked twin of _status */
    std::vector<int32_t> _status, _active;
    std::unique_ptr<WorkerPool> _pool;
    Eigen::MatrixXd _X;
    double _t_pack = 0.0;
    std::string _error;
};

}}

#endif

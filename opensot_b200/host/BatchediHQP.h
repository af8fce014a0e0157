/*
 * BatchediHQP.h -- batched front-end: B structurally identical OpenSoT stacks, one GPU solve.
 *
 * The batched counterpart of OpenSoT::solvers::iHQP (include/OpenSoT/solvers/iHQP.h:59-193).
 * It takes the SAME objects an existing program already has -- OpenSoT::AutoStack::Ptr built
 * with the Math-of-Tasks operators (+ * % / <<, include/OpenSoT/utils/AutoStack.h:115-336) --
 * one per robot instance / MPC sample, and replaces
 *     for each instance: iHQP::solve()         src/solvers/iHQP.cpp:263-358
 * by ONE call into the C ABI (include/opensot_b200.h). Per cycle the user still calls
 * AutoStack::update() on every stack (that is where the robot model produces A and b); what
 * this class replaces on the host is the piling: Aggregated::generateAll / MatrixPiler copies
 * (src/tasks/Aggregated.cpp:113-132, src/constraints/Aggregated.cpp:73-257,
 * include/OpenSoT/utils/Piler.h:102-163) become writes of every task's rows at fixed offsets
 * of the instance blob, and BackEnd::updateTask/updateConstraints/updateBounds deep copies
 * (src/solvers/BackEnd.cpp:19-104) disappear.
 *
 * Header-only on purpose: it is compiled against whatever OpenSoT headers the host program
 * uses, so the Task / Constraint / AutoStack ABI is the program's own.
 * Error behaviour mirrors the reference: constructors throw std::runtime_error
 * (iHQP.cpp:24), solve() returns false if any instance failed (iHQP.cpp:279,335-347) and
 * the per-instance osb_status words stay available through getStatus().
 */
#ifndef OPENSOT_B200_BATCHED_IHQP_H_
#define OPENSOT_B200_BATCHED_IHQP_H_

#include <OpenSoT/Constraint.h>
#include <OpenSoT/Task.h>
#include <OpenSoT/utils/AutoStack.h>

#include <cmath>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "opensot_b200.h"

namespace OpenSoT { namespace b200 {

class BatchediHQP {
public:
    typedef OpenSoT::Task<Eigen::MatrixXd, Eigen::VectorXd> TaskType;
    typedef OpenSoT::Constraint<Eigen::MatrixXd, Eigen::VectorXd> ConstraintType;
    typedef std::shared_ptr<TaskType> TaskPtr;
    typedef std::shared_ptr<ConstraintType> ConstraintPtr;

    /* eps_regularisation is the same FACTOR iHQP takes (default 2e2, iHQP.h:32) */
    BatchediHQP(const std::vector<OpenSoT::AutoStack::Ptr>& stacks, const double eps_regularisation = 2e2,
                const int device = 0)
        : _stacks(stacks), _engine(nullptr), _eps_factor(eps_regularisation)
    {
        if (_stacks.empty()) throw std::runtime_error("BatchediHQP: no stacks given");
        if (!prepare(device)) throw std::runtime_error("Can Not initizalize batched SoT: " + _error);
    }
    ~BatchediHQP() { osb_engine_destroy(_engine); }
    BatchediHQP(const BatchediHQP&) = delete;
    BatchediHQP& operator=(const BatchediHQP&) = delete;

    /* solutions: n x B, column i = joint velocities of instance i (iHQP::solve for all) */
    bool solve(Eigen::MatrixXd& solutions)
    {
        const int B = (int)_stacks.size(), n = _desc.n;
        for (int i = 0; i < B; ++i)
            if (!pack(i, _blobs.data() + (size_t)i * _layout.stride)) return false;
        if (solutions.rows() != n || solutions.cols() != B) solutions.resize(n, B);
        /* an n x B column-major Eigen matrix is B consecutive solution vectors */
        if (osb_solve_host(_engine, _blobs.data(), B, solutions.data(), _status.data(), nullptr) != OSB_SUCCESS) {
            _error = osb_last_error();
            return false;
        }
        for (int i = 0; i < B; ++i)
            if (_status[i] != OSB_OK) return false;
        return true;
    }
    /* AutoStack::update() on every instance (src/utils/AutoStack.cpp:385-393), then solve */
    bool updateAndSolve(Eigen::MatrixXd& solutions)
    {
        for (auto& s : _stacks) s->update();
        return solve(solutions);
    }

    unsigned int getNumberOfTasks() const { return (unsigned int)_desc.n_levels; }
    unsigned int getBatchSize() const { return (unsigned int)_stacks.size(); }
    void setActiveStack(const unsigned int i, const bool flag)
    {
        if (i < _active.size()) { _active[i] = flag ? 1 : 0; push_active(); }
    }
    void activateAllStacks() { _active.assign(_active.size(), 1); push_active(); }
    std::string getBackEndName(const unsigned int) const { return "B200"; }
    const std::vector<int32_t>& getStatus() const { return _status; }
    const std::string& getLastError() const { return _error; }
    const osb_stack_desc& getDescriptor() const { return _desc; }

private:
    static bool is_zero(const Eigen::VectorXd& v) { return v.size() == 0 || v.cwiseAbs().maxCoeff() == 0.0; }

    /* rows a list of constraints contributes: bounds become identity rows when `bounds_as_rows` */
    struct RowCount { int eq = 0, in = 0, bounds = 0; };

    bool prepare(const int device)
    {
        OpenSoT::AutoStack& s0 = *_stacks[0];
        auto& levels = s0.getStack();
        if (levels.empty() || levels.size() > OSB_MAX_LEVELS) { _error = "unsupported number of levels"; return false; }
        std::memset(&_desc, 0, sizeof(_desc));
        const int n = (int)levels[0]->getXSize();
        _desc.n = n;
        _desc.n_levels = (int)levels.size();
        _desc.eps = _eps_factor * OSB_QPOASES_BASE_EPS;
        for (auto& s : _stacks) s->update();
        for (size_t k = 0; k < levels.size(); ++k) {
            TaskPtr t = levels[k];
            _desc.m[k] = (int)t->getA().rows();
            _desc.task_kind[k] = OSB_TASK_DENSE;
            bool diag = true, ident = true, has_c = false;
            for (auto& s : _stacks) {
                TaskPtr tk = s->getStack()[k];
                if ((int)tk->getA().rows() != _desc.m[k] || (int)tk->getXSize() != n) { _error = "stacks differ in shape"; return false; }
                const Eigen::MatrixXd& W = tk->getWeight();
                Eigen::MatrixXd off = W; off.diagonal().setZero();
                if (off.cwiseAbs().maxCoeff() != 0.0) diag = false;
                if (!W.isIdentity(0.0)) ident = false;
                if (!is_zero(tk->getc())) has_c = true;
            }
            _desc.weight_kind[k] = ident ? OSB_W_IDENTITY : (diag ? OSB_W_DIAG : OSB_W_DENSE);
            _desc.has_c[k] = has_c ? 1 : 0;
            int rows = 0;
            for (auto& c : t->getConstraints()) rows += local_rows(*c, n);
            _desc.c_loc[k] = rows;
        }
        /* global constraints and bounds: what `stack << ...` collected (AutoStack.cpp:303-316) */
        _n_bound_cons = 0;
        int rows = 0;
        for (auto& c : s0.getBoundsList()) {
            if (c->isBound()) ++_n_bound_cons;
            else rows += local_rows(*c, n);
        }
        _desc.n_bound_sets = _n_bound_cons > 0 ? 1 : 0; /* intersected on the host while packing */
        /* equality rows are structural: a global row is an equality if bL == bU on every instance */
        _global_rows = rows;
        _row_is_eq.assign((size_t)rows, 1);
        std::vector<double> C((size_t)rows * n), lo((size_t)rows), hi((size_t)rows);
        for (auto& s : _stacks) {
            gather_rows(s->getBoundsList(), n, false, C.data(), lo.data(), hi.data());
            for (int r = 0; r < rows; ++r)
                if (!(lo[r] == hi[r])) _row_is_eq[r] = 0;
        }
        for (int r = 0; r < rows; ++r) (_row_is_eq[r] ? _desc.c_eq : _desc.c_in)++;
        TaskPtr reg = s0.getRegularisationTask();
        if (reg) {
            _desc.m_reg = (int)reg->getA().rows();
            _desc.reg_weight_kind = reg->getWeight().isIdentity(0.0) ? OSB_W_IDENTITY : OSB_W_DIAG;
        }
        if (osb_blob_layout(&_desc, &_layout) != OSB_SUCCESS) { _error = "stack shape not supported by the engine"; return false; }
        const int B = (int)_stacks.size();
        if (osb_engine_create(&_engine, &_desc, B, device) != OSB_SUCCESS) { _error = osb_last_error(); return false; }
        _blobs.assign((size_t)B * _layout.stride, 0.0);
        _status.assign((size_t)B, 0);
        _active.assign((size_t)_desc.n_levels, 1);
        _tmpC.resize((size_t)(rows > 0 ? rows : 1) * n); _tmplo.resize(rows + 1); _tmphi.resize(rows + 1);
        return true;
    }

    static int local_rows(ConstraintType& c, const int n)
    {
        int rows = 0;
        if (c.getAineq().rows() > 0) rows += (int)c.getAineq().rows();
        if (c.getAeq().rows() > 0) rows += (int)c.getAeq().rows();
        if (c.getAineq().rows() == 0 && c.getAeq().rows() == 0 && c.hasBounds()) rows += n; /* local bound */
        return rows;
    }

    /* writes the rows of every non-bound constraint of the list (and of bounds too when
     * bounds_as_rows) in list order: what constraints::Aggregated::generateAll piles into
     * Aineq / bLowerBound / bUpperBound (src/constraints/Aggregated.cpp:159-203) */
    template <class List>
    static int gather_rows(List& list, const int n, const bool bounds_as_rows, double* C, double* lo, double* hi)
    {
        const double inf = std::numeric_limits<double>::infinity();
        int r = 0;
        for (auto& cp : list) {
            ConstraintType& c = *cp;
            const bool is_pure_bound = (c.getAineq().rows() == 0 && c.getAeq().rows() == 0 && c.hasBounds());
            if (is_pure_bound) {
                if (!bounds_as_rows) continue;
                for (int i = 0; i < n; ++i, ++r) {
                    for (int j = 0; j < n; ++j) C[(size_t)r * n + j] = (i == j) ? 1.0 : 0.0;
                    lo[r] = c.getLowerBound().size() ? c.getLowerBound()[i] : -inf;
                    hi[r] = c.getUpperBound().size() ? c.getUpperBound()[i] : inf;
                }
                continue;
            }
            const Eigen::MatrixXd& Aeq = c.getAeq();
            for (int i = 0; i < Aeq.rows(); ++i, ++r) { /* Aeq x = beq  ->  beq <= Aeq x <= beq */
                for (int j = 0; j < n; ++j) C[(size_t)r * n + j] = Aeq(i, j);
                lo[r] = hi[r] = c.getbeq()[i];
            }
            const Eigen::MatrixXd& Ai = c.getAineq();
            for (int i = 0; i < Ai.rows(); ++i, ++r) { /* one-sided rows get an infinite other side */
                for (int j = 0; j < n; ++j) C[(size_t)r * n + j] = Ai(i, j);
                lo[r] = c.getbLowerBound().size() ? c.getbLowerBound()[i] : -inf;
                hi[r] = c.getbUpperBound().size() ? c.getbUpperBound()[i] : inf;
            }
        }
        return r;
    }

    bool pack(const int inst, double* blob)
    {
        OpenSoT::AutoStack& s = *_stacks[inst];
        const int n = _desc.n;
        auto& levels = s.getStack();
        for (int k = 0; k < _desc.n_levels; ++k) {
            TaskType& t = *levels[k];
            const int m = _desc.m[k];
            const Eigen::MatrixXd& A = t.getA();
            if (A.rows() != m || A.cols() != n) { _error = "task changed size"; return false; }
            double* dst = blob + _layout.A[k];
            for (int r = 0; r < m; ++r)
                for (int j = 0; j < n; ++j) dst[r * n + j] = A(r, j);
            std::memcpy(blob + _layout.b[k], t.getb().data(), sizeof(double) * m);
            const Eigen::MatrixXd& W = t.getWeight();
            if (_desc.weight_kind[k] == OSB_W_DIAG)
                for (int r = 0; r < m; ++r) blob[_layout.w[k] + r] = W(r, r);
            else if (_desc.weight_kind[k] == OSB_W_DENSE)
                for (int r = 0; r < m; ++r)
                    for (int j = 0; j < m; ++j) blob[_layout.w[k] + r * m + j] = W(r, j);
            else if (!W.isIdentity(0.0)) { _error = "weight of level " + std::to_string(k) + " stopped being the identity"; return false; }
            if (_desc.has_c[k]) std::memcpy(blob + _layout.c[k], t.getc().data(), sizeof(double) * n);
            else if (!is_zero(t.getc())) { _error = "linear cost term appeared after construction"; return false; }
            if (_desc.c_loc[k] > 0) {
                const int got = gather_rows(t.getConstraints(), n, true, blob + _layout.Cl[k],
                                            blob + _layout.lCl[k], blob + _layout.uCl[k]);
                if (got != _desc.c_loc[k]) { _error = "task-local constraint rows changed"; return false; }
            }
        }
        /* bounds: l = max l_i, u = min u_i over every bound of the stack (Aggregated.cpp:141-148) */
        if (_n_bound_cons > 0) {
            double* l = blob + _layout.l[0];
            double* u = blob + _layout.u[0];
            for (int i = 0; i < n; ++i) { l[i] = -std::numeric_limits<double>::infinity(); u[i] = -l[i]; }
            for (auto& c : s.getBoundsList()) {
                if (!c->isBound()) continue;
                for (int i = 0; i < n; ++i) {
                    if (c->getLowerBound().size()) l[i] = std::max(l[i], c->getLowerBound()[i]);
                    if (c->getUpperBound().size()) u[i] = std::min(u[i], c->getUpperBound()[i]);
                }
            }
        }
        if (_global_rows > 0) {
            const int got = gather_rows(s.getBoundsList(), n, false, _tmpC.data(), _tmplo.data(), _tmphi.data());
            if (got != _global_rows) { _error = "global constraint rows changed"; return false; }
            int e = 0, q = 0;
            for (int r = 0; r < _global_rows; ++r) {
                if (_row_is_eq[r]) {
                    if (!(_tmplo[r] == _tmphi[r])) { _error = "an equality row became an inequality"; return false; }
                    std::memcpy(blob + _layout.Ceq + (size_t)e * n, _tmpC.data() + (size_t)r * n, sizeof(double) * n);
                    blob[_layout.deq + e] = _tmplo[r];
                    ++e;
                } else {
                    std::memcpy(blob + _layout.Cin + (size_t)q * n, _tmpC.data() + (size_t)r * n, sizeof(double) * n);
                    blob[_layout.lA + q] = _tmplo[r];
                    blob[_layout.uA + q] = _tmphi[r];
                    ++q;
                }
            }
        }
        if (_desc.m_reg > 0) {
            TaskPtr reg = s.getRegularisationTask();
            const int mr = _desc.m_reg;
            for (int r = 0; r < mr; ++r)
                for (int j = 0; j < n; ++j) blob[_layout.Ar + r * n + j] = reg->getA()(r, j);
            std::memcpy(blob + _layout.br, reg->getb().data(), sizeof(double) * mr);
            if (_desc.reg_weight_kind == OSB_W_DIAG)
                for (int r = 0; r < mr; ++r) blob[_layout.wr + r] = reg->getWeight()(r, r);
        }
        return true;
    }

    void push_active() { osb_set_active_levels(_engine, _active.data()); }

    std::vector<OpenSoT::AutoStack::Ptr> _stacks;
    osb_engine* _engine;
    osb_stack_desc _desc;
    osb_layout _layout;
    double _eps_factor;
    int _n_bound_cons = 0, _global_rows = 0;
    std::vector<char> _row_is_eq;
    std::vector<double> _blobs, _tmpC, _tmplo, _tmphi;
    std::vector<int32_t> _status, _active;
    std::string _error;
};

}}

#endif

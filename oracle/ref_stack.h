// ref_stack.h -- TEST INFRASTRUCTURE: builds an OpenSoT AutoStack (reference API, Math of
// Tasks operators) from a synthetic instance blob. Shared by oracle/ref_driver.cpp and
// tests/cpp/test_dropin.cpp. Never included by product code.
#ifndef OSB_REF_STACK_H_
#define OSB_REF_STACK_H_
#include <OpenSoT/constraints/GenericConstraint.h>
#include <OpenSoT/constraints/TaskToConstraint.h>
#include <OpenSoT/solvers/iHQP.h>
#include <OpenSoT/tasks/GenericTask.h>
#include <OpenSoT/utils/AutoStack.h>
#include <memory>
#include <string>
#include <vector>
#include "opensot_b200.h"

namespace refstack {
using namespace OpenSoT;
typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> RowMat;

struct RefStack {
    std::vector<tasks::GenericTask::Ptr> level_task;
    std::vector<constraints::GenericConstraint::Ptr> local_con;
    tasks::GenericTask::Ptr feet, reg;
    constraints::GenericConstraint::Ptr global_con;
    std::vector<constraints::GenericConstraint::Ptr> bounds;
    AutoStack::Ptr stack;
    std::shared_ptr<solvers::iHQP> solver;
};

static Eigen::MatrixXd weight_of(const osb_stack_desc& d, const osb_layout& L, int k,
                                 const double* blob)
{
    const int m = d.m[k];
    Eigen::MatrixXd W = Eigen::MatrixXd::Identity(m, m);
    if (d.weight_kind[k] == OSB_W_DIAG)
        for (int i = 0; i < m; ++i) W(i, i) = blob[L.w[k] + i];
    else if (d.weight_kind[k] == OSB_W_DENSE)
        W = Eigen::Map<const RowMat>(blob + L.w[k], m, m);
    return W;
}

// Push the numbers of one blob into already constructed reference objects (what a user's
// control loop does every cycle before stack->update()).
static void set_data(RefStack& rs, const osb_stack_desc& d, const osb_layout& L,
                     const double* blob)
{
    const int n = d.n;
    for (int k = 0; k < d.n_levels; ++k) {
        const int m = d.m[k];
        Eigen::VectorXd b = Eigen::Map<const Eigen::VectorXd>(blob + L.b[k], m);
        if (d.task_kind[k] == OSB_TASK_DENSE) {
            Eigen::MatrixXd A = Eigen::Map<const RowMat>(blob + L.A[k], m, n);
            rs.level_task[k]->setA(A);
        }
        rs.level_task[k]->setb(b);
        if (d.weight_kind[k] != OSB_W_IDENTITY)
            rs.level_task[k]->setWeight(weight_of(d, L, k, blob));
        if (d.has_c[k])
            rs.level_task[k]->setc(Eigen::Map<const Eigen::VectorXd>(blob + L.c[k], n));
        if (d.c_loc[k] > 0) {
            Eigen::MatrixXd C = Eigen::Map<const RowMat>(blob + L.Cl[k], d.c_loc[k], n);
            AffineHelper var(C, Eigen::VectorXd::Zero(d.c_loc[k]));
            rs.local_con[k]->setConstraint(
                var, Eigen::Map<const Eigen::VectorXd>(blob + L.uCl[k], d.c_loc[k]),
                Eigen::Map<const Eigen::VectorXd>(blob + L.lCl[k], d.c_loc[k]));
        }
    }
    if (d.c_eq > 0) {
        Eigen::MatrixXd C = Eigen::Map<const RowMat>(blob + L.Ceq, d.c_eq, n);
        rs.feet->setA(C);
        rs.feet->setb(Eigen::Map<const Eigen::VectorXd>(blob + L.deq, d.c_eq));
    }
    if (d.c_in > 0) {
        Eigen::MatrixXd C = Eigen::Map<const RowMat>(blob + L.Cin, d.c_in, n);
        AffineHelper var(C, Eigen::VectorXd::Zero(d.c_in));
        rs.global_con->setConstraint(var,
                                     Eigen::Map<const Eigen::VectorXd>(blob + L.uA, d.c_in),
                                     Eigen::Map<const Eigen::VectorXd>(blob + L.lA, d.c_in));
    }
    for (int s = 0; s < d.n_bound_sets; ++s)
        rs.bounds[s]->setBounds(Eigen::Map<const Eigen::VectorXd>(blob + L.u[s], n),
                                Eigen::Map<const Eigen::VectorXd>(blob + L.l[s], n));
    if (d.m_reg > 0) {
        Eigen::MatrixXd A = Eigen::Map<const RowMat>(blob + L.Ar, d.m_reg, n);
        rs.reg->setA(A);
        rs.reg->setb(Eigen::Map<const Eigen::VectorXd>(blob + L.br, d.m_reg));
        if (d.reg_weight_kind == OSB_W_DIAG) {
            Eigen::MatrixXd W = Eigen::MatrixXd::Identity(d.m_reg, d.m_reg);
            for (int i = 0; i < d.m_reg; ++i) W(i, i) = blob[L.wr + i];
            rs.reg->setWeight(W);
        }
    }
}

static void build(RefStack& rs, const osb_stack_desc& d, const osb_layout& L,
                  const double* blob, solvers::solver_back_ends be, bool with_solver = true)
{
    const int n = d.n;
    rs.level_task.resize(d.n_levels);
    rs.local_con.resize(d.n_levels);
    for (int k = 0; k < d.n_levels; ++k) {
        const int m = d.m[k];
        Eigen::MatrixXd A = (d.task_kind[k] == OSB_TASK_DENSE)
                                ? Eigen::MatrixXd(Eigen::Map<const RowMat>(blob + L.A[k], m, n))
                                : Eigen::MatrixXd(Eigen::MatrixXd::Identity(n, n));
        Eigen::VectorXd b = Eigen::Map<const Eigen::VectorXd>(blob + L.b[k], m);
        rs.level_task[k] =
            std::make_shared<tasks::GenericTask>("t" + std::to_string(k), A, b);
        if (d.c_loc[k] > 0) {
            Eigen::MatrixXd C = Eigen::Map<const RowMat>(blob + L.Cl[k], d.c_loc[k], n);
            AffineHelper var(C, Eigen::VectorXd::Zero(d.c_loc[k]));
            rs.local_con[k] = std::make_shared<constraints::GenericConstraint>(
                "lc" + std::to_string(k), var,
                Eigen::Map<const Eigen::VectorXd>(blob + L.uCl[k], d.c_loc[k]),
                Eigen::Map<const Eigen::VectorXd>(blob + L.lCl[k], d.c_loc[k]),
                constraints::GenericConstraint::Type::CONSTRAINT);
            tasks::Aggregated::TaskPtr tp = rs.level_task[k];
            tp << rs.local_con[k];
        }
    }
    tasks::Aggregated::TaskPtr t0 = rs.level_task[0];
    if (d.n_levels == 1) {
        rs.stack = std::make_shared<AutoStack>(t0);
    } else {
        tasks::Aggregated::TaskPtr t1 = rs.level_task[1];
        rs.stack = t0 / t1;
        for (int k = 2; k < d.n_levels; ++k) {
            tasks::Aggregated::TaskPtr tk = rs.level_task[k];
            rs.stack = rs.stack / tk;
        }
    }
    rs.bounds.resize(d.n_bound_sets);
    for (int s = 0; s < d.n_bound_sets; ++s) {
        rs.bounds[s] = std::make_shared<constraints::GenericConstraint>(
            "bound" + std::to_string(s), Eigen::Map<const Eigen::VectorXd>(blob + L.u[s], n),
            Eigen::Map<const Eigen::VectorXd>(blob + L.l[s], n), n);
        rs.stack = rs.stack << rs.bounds[s];
    }
    if (d.c_in > 0) {
        Eigen::MatrixXd C = Eigen::Map<const RowMat>(blob + L.Cin, d.c_in, n);
        AffineHelper var(C, Eigen::VectorXd::Zero(d.c_in));
        rs.global_con = std::make_shared<constraints::GenericConstraint>(
            "gc", var, Eigen::Map<const Eigen::VectorXd>(blob + L.uA, d.c_in),
            Eigen::Map<const Eigen::VectorXd>(blob + L.lA, d.c_in),
            constraints::GenericConstraint::Type::CONSTRAINT);
        rs.stack = rs.stack << rs.global_con;
    }
    if (d.c_eq > 0) {
        Eigen::MatrixXd C = Eigen::Map<const RowMat>(blob + L.Ceq, d.c_eq, n);
        Eigen::VectorXd de = Eigen::Map<const Eigen::VectorXd>(blob + L.deq, d.c_eq);
        rs.feet = std::make_shared<tasks::GenericTask>("feet", C, de);
        tasks::Aggregated::TaskPtr fp = rs.feet;
        rs.stack = rs.stack << fp;
    }
    if (d.m_reg > 0) {
        Eigen::MatrixXd A = Eigen::Map<const RowMat>(blob + L.Ar, d.m_reg, n);
        Eigen::VectorXd b = Eigen::Map<const Eigen::VectorXd>(blob + L.br, d.m_reg);
        rs.reg = std::make_shared<tasks::GenericTask>("reg", A, b);
        rs.stack->setRegularisationTask(rs.reg);
    }
    set_data(rs, d, L, blob);
    rs.stack->update();
    if (!with_solver) return;
    const double eps_factor = d.eps / OSB_QPOASES_BASE_EPS;
    rs.solver = std::make_shared<solvers::iHQP>(*rs.stack, eps_factor, be);
}

} // namespace refstack
#endif

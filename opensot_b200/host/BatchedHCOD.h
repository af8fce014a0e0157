/*
 * BatchedHCOD.h -- the HCOD front-end surface over the same GPU engine.
 *
 * OpenSoT::solvers::HCOD (include/OpenSoT/solvers/HCOD.h, src/solvers/HCOD.cpp:107-219) wraps
 * soth's hierarchical complete orthogonal decomposition: a true lexicographic solver in which
 * bounds and constraints form the first stage. BASELINE.json names it among the replaced
 * subsystems; its 13-21 ms per solve (docs/source/data/coman_ik_stats_3_LEVELS_hcod.mat) is the
 * slowest reference path. The cascade the B200 engine solves has the same lexicographic optimum
 * in the limit eps -> 0 (measured agreement of soth with iHQP+qpOASES: median 7e-13,
 * SURVEY.md section 8c), so this class keeps HCOD's constructors (HCOD.h:33-41: AutoStack + damping,
 * Stack + bounds + damping), solve / setDamping / setDisableWeightsComputation surface and forwards
 * to BatchediHQP: it is a SURFACE over the same engine, not an implementation of the HCOD algorithm.
 * `damping` plays the role of soth's per-stage Tikhonov term: the engine's diagonal shift is
 * eps = max(damping^2, iHQP's default 2e2 * 2.22e-13).
 */
#ifndef OPENSOT_B200_BATCHED_HCOD_H_
#define OPENSOT_B200_BATCHED_HCOD_H_

#include "BatchediHQP.h"

namespace OpenSoT { namespace b200 {

class BatchedHCOD : public OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> {
public:
    typedef std::shared_ptr<BatchedHCOD> Ptr;
    typedef OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> Base;

    /* the batch: one AutoStack per instance */
    BatchedHCOD(const std::vector<OpenSoT::AutoStack::Ptr>& stacks, const double damping, const int device = 0)
        : Base(stacks.at(0)->getStack(), stacks.at(0)->getBounds()), _damping(damping)
    {
        _impl.reset(new BatchediHQP(stacks, factor_of(damping), device));
    }
    /* HCOD(AutoStack&, damping), HCOD.h:33 */
    BatchedHCOD(OpenSoT::AutoStack& stack_of_tasks, const double damping)
        : Base(stack_of_tasks.getStack(), stack_of_tasks.getBounds()), _damping(damping)
    {
        _impl.reset(new BatchediHQP(stack_of_tasks, factor_of(damping)));
    }
    /* HCOD(Stack&, ConstraintPtr bounds, damping), HCOD.h:41 */
    BatchedHCOD(Base::Stack& stack_of_tasks, Base::ConstraintPtr bounds, const double damping)
        : Base(stack_of_tasks, bounds), _damping(damping)
    {
        _impl.reset(new BatchediHQP(stack_of_tasks, bounds, factor_of(damping)));
    }
    /* the batch, Stack + bounds per instance */
    BatchedHCOD(std::vector<Base::Stack>& stacks, const std::vector<Base::ConstraintPtr>& bounds, const double damping,
                const int device = 0)
        : Base(stacks.at(0), bounds.at(0)), _damping(damping)
    {
        _impl.reset(new BatchediHQP(stacks, bounds, factor_of(damping), device));
    }

    /* n x B solutions, like BatchediHQP::solve; AutoStack::update() is the caller's job, as
     * with HCOD::solve (src/solvers/HCOD.cpp:107-124) */
    bool solve(Eigen::MatrixXd& solutions) { return _impl->solve(solutions); }
    bool solve(Eigen::VectorXd& solution) override { return _impl->solve(solution); }
    /* HCOD::setDamping (HCOD.h:77-81): to all levels */
    void setDamping(double damping)
    {
        _damping = damping;
        _impl->setEpsRegularisation(factor_of(damping) * OSB_QPOASES_BASE_EPS);
    }
    double getDamping() const { return _damping; }
    /* soth needs W^(1/2) per cycle (HCOD.cpp:154-158) and offers to skip it; weights enter the
     * engine as A'WA directly, so there is nothing to disable: kept for source compatibility */
    void setDisableWeightsComputation(const bool disable) { _disable_weights = disable; }
    bool getDisableWeightsComputation() { return _disable_weights; }
    const std::vector<int32_t>& getStatus() const { return _impl->getStatus(); }
    BatchediHQP& engine() { return *_impl; }

private:
    static double factor_of(const double damping)
    {
        const double eps_abs = std::max(damping * damping, 2e2 * OSB_QPOASES_BASE_EPS);
        return eps_abs / OSB_QPOASES_BASE_EPS;
    }
    std::unique_ptr<BatchediHQP> _impl;
    double _damping;
    bool _disable_weights = false;
};

}}

#endif

/*
 * BatchedHCOD.h -- the HCOD front-end surface over the same GPU engine.
 *
 * OpenSoT::solvers::HCOD (include/OpenSoT/solvers/HCOD.h, src/solvers/HCOD.cpp:107-219) wraps
 * soth's hierarchical complete orthogonal decomposition: a true lexicographic solver in which
 * bounds and constraints form the first stage. BASELINE.json names it among the replaced
 * subsystems; its 13-21 ms per solve (docs/source/data/coman_ik_stats_3_LEVELS_hcod.mat) is the
 * slowest reference path. The cascade the B200 engine solves has the same lexicographic optimum
 * in the limit eps -> 0 (measured agreement of soth with iHQP+qpOASES: median 7e-13,
 * SURVEY.md section 8c), so this class keeps HCOD's constructor / solve / setDamping surface
 * and forwards to BatchediHQP. `damping` plays the role of soth's per-stage Tikhonov term: the
 * engine's diagonal shift is eps = max(damping^2, iHQP's default 2e2 * 2.22e-13).
 */
#ifndef OPENSOT_B200_BATCHED_HCOD_H_
#define OPENSOT_B200_BATCHED_HCOD_H_

#include "BatchediHQP.h"

namespace OpenSoT { namespace b200 {

class BatchedHCOD {
public:
    BatchedHCOD(const std::vector<OpenSoT::AutoStack::Ptr>& stacks, const double damping, const int device = 0)
        : _stacks(stacks), _device(device), _disable_weights(false)
    {
        setDamping(damping);
    }
    /* n x B solutions, like BatchediHQP::solve; AutoStack::update() is the caller's job, as
     * with HCOD::solve (src/solvers/HCOD.cpp:107-124) */
    bool solve(Eigen::MatrixXd& solutions) { return _impl->solve(solutions); }
    void setDamping(double damping)
    {
        _damping = damping;
        const double eps_abs = std::max(damping * damping, 2e2 * OSB_QPOASES_BASE_EPS);
        _impl.reset(new BatchediHQP(_stacks, eps_abs / OSB_QPOASES_BASE_EPS, _device));
    }
    double getDamping() const { return _damping; }
    /* soth needs W^(1/2) per cycle (HCOD.cpp:154-158) and offers to skip it; weights enter the
     * engine as A'WA directly, so there is nothing to disable: kept for source compatibility */
    void setDisableWeightsComputation(const bool disable) { _disable_weights = disable; }
    bool getDisableWeightsComputation() const { return _disable_weights; }
    const std::vector<int32_t>& getStatus() const { return _impl->getStatus(); }

private:
    std::vector<OpenSoT::AutoStack::Ptr> _stacks;
    std::unique_ptr<BatchediHQP> _impl;
    double _damping;
    int _device;
    bool _disable_weights;
};

}}

#endif

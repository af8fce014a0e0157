/*
 * B200BackEnd.h -- single-instance OpenSoT back-end plugin over the B200 engine.
 *
 * Drop-in for the reference's back-end plugins (src/solvers/QPOasesBackEnd.cpp:14-24):
 * built as libOpenSotBackEndODYS.so it is what `BackEndFactory(solver_back_ends::ODYS, ...)`
 * dlopens (the ODYS slot of the closed enum include/OpenSoT/solvers/BackEndFactory.h:10-18
 * has no implementation in the reference tree), so an unmodified `iHQP` can drive the GPU
 * solver level by level. This is the batch = 1 compatibility path (BASELINE cfg 1); the
 * throughput path is OpenSoT::b200::BatchediHQP (BatchediHQP.h).
 *
 * Host side only: Eigen types at the boundary, every number crosses into CUDA through the
 * C ABI of include/opensot_b200.h. There is no CPU solve in here.
 */
#ifndef OPENSOT_B200_BACKEND_H_
#define OPENSOT_B200_BACKEND_H_

#include <OpenSoT/solvers/BackEnd.h>
#include <vector>
#include "opensot_b200.h"

namespace OpenSoT { namespace solvers {

class B200BackEnd : public BackEnd {
public:
    /* same arguments as every reference plugin: QPOasesBackEnd.cpp:26-46 */
    B200BackEnd(const int number_of_variables, const int number_of_constraints,
                OpenSoT::HessianType hessian_type, const double eps_regularisation);
    ~B200BackEnd();

    bool initProblem(const Eigen::MatrixXd& H, const Eigen::VectorXd& g, const Eigen::MatrixXd& A,
                     const Eigen::VectorXd& lA, const Eigen::VectorXd& uA, const Eigen::VectorXd& l,
                     const Eigen::VectorXd& u) override;
    bool updateTask(const Eigen::MatrixXd& H, const Eigen::VectorXd& g) override;
    bool updateConstraints(const Eigen::Ref<const Eigen::MatrixXd>& A,
                           const Eigen::Ref<const Eigen::VectorXd>& lA,
                           const Eigen::Ref<const Eigen::VectorXd>& uA) override;
    bool updateBounds(const Eigen::VectorXd& l, const Eigen::VectorXd& u) override;
    bool solve() override;

    boost::any getOptions() override;              /* osb_options                         */
    void setOptions(const boost::any& options) override;
    double getObjective() override;                /* 1/2 x'Hx + g'x at the solution      */
    bool setEpsRegularisation(const double eps) override;   /* absolute, as QPOasesBackEnd */
    double getEpsRegularisation() override;

    int getLastStatus() const { return _status; }  /* osb_status word of the last solve   */

private:
    bool rebuild();                                /* (re)creates the engine for the current sizes */
    void _printProblemInformation() override;

    osb_engine* _engine;
    osb_stack_desc _desc;
    osb_layout _layout;
    osb_options _opt;
    std::vector<double> _blob;
    double _eps;                                   /* absolute diagonal shift              */
    int _status;
    bool _has_bounds;
};

}}

#endif

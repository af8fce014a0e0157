/* B200BackEnd.cpp -- see B200BackEnd.h. */
#include "B200BackEnd.h"

#include <cstring>
#include <stdexcept>

using namespace OpenSoT::solvers;

/* the symbols BackEndFactory looks up (src/solvers/BackEndFactory.cpp:4-17) */
extern "C" BackEnd* create_instance(const int number_of_variables, const int number_of_constraints,
                                    OpenSoT::HessianType hessian_type, const double eps_regularisation)
{
    return new B200BackEnd(number_of_variables, number_of_constraints, hessian_type, eps_regularisation);
}
extern "C" void destroy_instance(BackEnd* instance) { delete instance; }

B200BackEnd::B200BackEnd(const int number_of_variables, const int number_of_constraints,
                         OpenSoT::HessianType, const double eps_regularisation)
    : BackEnd(number_of_variables, number_of_constraints),
      _engine(nullptr),
      /* factor times qpOASES' base epsRegularisation, QPOasesBackEnd.cpp:57,67 */
      _eps(eps_regularisation * OSB_QPOASES_BASE_EPS),
      _status(OSB_OK),
      _has_bounds(true)
{
    std::memset(&_desc, 0, sizeof(_desc));
    std::memset(&_opt, 0, sizeof(_opt));
    _opt.tol_feas = 1e-12;
    _opt.tol_rank = 1e-11;
}

B200BackEnd::~B200BackEnd() { osb_engine_destroy(_engine); }

bool B200BackEnd::rebuild()
{
    osb_engine_destroy(_engine);
    _engine = nullptr;
    std::memset(&_desc, 0, sizeof(_desc));
    _desc.n = (int32_t)_H.rows();
    _desc.n_levels = 1;
    _desc.task_kind[0] = OSB_TASK_HESSIAN;
    _desc.weight_kind[0] = OSB_W_IDENTITY;
    _desc.c_in = (int32_t)_A.rows();
    _desc.n_bound_sets = _has_bounds ? 1 : 0;
    _desc.eps = _eps;
    if (osb_blob_layout(&_desc, &_layout) != OSB_SUCCESS) {
        XBot::Logger::error("B200BackEnd: unsupported problem size (%i variables, %i constraints)\n",
                            (int)_H.rows(), (int)_A.rows());
        return false;
    }
    if (osb_engine_create(&_engine, &_desc, 1, 0) != OSB_SUCCESS) {
        XBot::Logger::error("B200BackEnd: %s\n", osb_last_error());
        return false;
    }
    osb_set_options(_engine, &_opt);
    _blob.assign((size_t)_layout.stride, 0.0);
    return true;
}

bool B200BackEnd::initProblem(const Eigen::MatrixXd& H, const Eigen::VectorXd& g, const Eigen::MatrixXd& A,
                              const Eigen::VectorXd& lA, const Eigen::VectorXd& uA,
                              const Eigen::VectorXd& l, const Eigen::VectorXd& u)
{
    /* same checks as QPOasesBackEnd::initProblem (QPOasesBackEnd.cpp:86-115) */
    if (A.rows() != _A.rows()) {
        XBot::Logger::error("A.rows() != _A.rows() --> %i != %i", (int)A.rows(), (int)_A.rows());
        return false;
    }
    if (l.rows() != u.rows() || lA.rows() != A.rows() || lA.rows() != uA.rows() || H.rows() != H.cols() ||
        g.rows() != H.rows() || (A.rows() > 0 && A.cols() != H.cols()))
        return false;
    _H = H; _g = g; _A = A; _lA = lA; _uA = uA; _l = l; _u = u;
    _has_bounds = (l.rows() > 0);
    if (_has_bounds && l.rows() != H.rows()) return false;
    if (!rebuild()) return false;
    return solve();
}

bool B200BackEnd::updateTask(const Eigen::MatrixXd& H, const Eigen::VectorXd& g)
{
    if (g.size() != _g.size() || H.cols() != _H.cols()) {
        if (H.rows() != H.cols() || g.size() != H.rows()) return false;
        _H = H; _g = g; /* size change: the reference rebuilds its QP object, :170-201 */
        return rebuild();
    }
    _H = H; _g = g;
    return true;
}

bool B200BackEnd::updateConstraints(const Eigen::Ref<const Eigen::MatrixXd>& A,
                                    const Eigen::Ref<const Eigen::VectorXd>& lA,
                                    const Eigen::Ref<const Eigen::VectorXd>& uA)
{
    if (A.rows() > 0 && A.cols() != _H.cols()) return false;
    if (lA.rows() != A.rows() || lA.rows() != uA.rows()) return false;
    const bool resized = (A.rows() != _A.rows());
    _A = A; _lA = lA; _uA = uA;
    if (resized) return rebuild(); /* row count changed: QPOasesBackEnd.cpp:229-244 */
    return true;
}

bool B200BackEnd::updateBounds(const Eigen::VectorXd& l, const Eigen::VectorXd& u)
{
    if (l.rows() != u.rows()) return false;
    if (l.rows() != _l.rows()) {
        if (l.rows() != 0 && l.rows() != _H.rows()) return false;
        _l = l; _u = u; _has_bounds = (l.rows() > 0);
        return rebuild();
    }
    _l = l; _u = u;
    return true;
}

bool B200BackEnd::solve()
{
    if (!_engine && !rebuild()) return false;
    const int n = _desc.n, nc = _desc.c_in;
    double* b = _blob.data();
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) b[_layout.A[0] + i * n + j] = _H(i, j); /* eps is added on device */
    for (int i = 0; i < n; ++i) b[_layout.b[0] + i] = _g[i];
    for (int r = 0; r < nc; ++r) {
        for (int j = 0; j < n; ++j) b[_layout.Cin + r * n + j] = _A(r, j); /* row-major, as :257 */
        b[_layout.lA + r] = _lA[r];
        b[_layout.uA + r] = _uA[r];
    }
    if (_has_bounds)
        for (int i = 0; i < n; ++i) { b[_layout.l[0] + i] = _l[i]; b[_layout.u[0] + i] = _u[i]; }
    if (_solution.rows() != n) _solution.resize(n);
    int32_t st = 0;
    if (osb_solve_host(_engine, b, 1, _solution.data(), &st, nullptr) != OSB_SUCCESS) {
        XBot::Logger::error("B200BackEnd: %s\n", osb_last_error());
        return false;
    }
    _status = st;
    return st == OSB_OK; /* infeasible / iteration limit => false, like the reference */
}

boost::any B200BackEnd::getOptions() { return boost::any(_opt); }

void B200BackEnd::setOptions(const boost::any& options)
{
    _opt = boost::any_cast<osb_options>(options);
    if (_engine) osb_set_options(_engine, &_opt);
}

double B200BackEnd::getObjective()
{
    return 0.5 * _solution.dot(_H * _solution) + _g.dot(_solution);
}

bool B200BackEnd::setEpsRegularisation(const double eps)
{
    if (eps < 0.0) {
        XBot::Logger::error("Negative eps is not allowed!");
        return false;
    }
    _eps = eps;
    return _engine ? rebuild() : true;
}

double B200BackEnd::getEpsRegularisation() { return _eps; }

void B200BackEnd::_printProblemInformation()
{
    XBot::Logger::info("B200 back-end: %s, eps %e\n", osb_version(), _eps);
}

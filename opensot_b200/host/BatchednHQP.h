/*
 * BatchednHQP.h -- batched front-end with the semantics of OpenSoT::solvers::nHQP.
 *
 * The reference's nHQP (include/OpenSoT/solvers/nHQP.h, src/solvers/nHQP.cpp) solves the hierarchy in the
 * cumulated null space: level i optimises x_i in R^{f_i}, the joint velocities are q = q0 + N_i x_i,
 *     AN = A_i N_i, b0 = b_i - A_i q0                                    (nHQP.cpp:357-372)
 *     SVD of AN; small singular values and the matching parts of b0 are damped
 *     (regularize_A_b, nHQP.cpp:236-280: threshold = min_sv_ratio * sv_max)
 *     H = AN' W AN (+ sv_max Vn Vn' on the null space of the level), g = -AN' W b0   (nHQP.cpp:377-388)
 *     rows [C N_i ; N_i] with limits shifted by C q0 and q0: bounds stay hard at every level (nHQP.cpp:283-317)
 *     one QP per level through a BackEnd (H_ii += eps inside it), q0 += N_i x_i, N_{i+1} = N_i Vn   (nHQP.cpp:155-204)
 * Here the B instances go through that recursion together: the per-level preparation (the SVD with its
 * regularisation, H, g and the projected rows) runs on the host, threads over instances, with the same
 * Eigen decomposition class family the reference uses, and every level's B QPs are ONE call into the
 * C ABI: a single-level OSB_TASK_HESSIAN stack of f_i variables on the same kernels that solve iHQP levels
 * (include/opensot_b200.h). The layer sizes f_i come from the rank of the first instance at construction
 * (nHQP.cpp:93-113, SV_THRESH 1e-6) and must be common to the batch.
 *
 * Same restrictions as the reference: no task-local constraints, no equality constraints in `bounds`
 * (nHQP.cpp:39-50); constructors throw std::runtime_error like the reference's.
 */
#ifndef OPENSOT_B200_BATCHED_NHQP_H_
#define OPENSOT_B200_BATCHED_NHQP_H_

#include <Eigen/SVD>

#include "BatchediHQP.h"

namespace OpenSoT { namespace b200 {

class BatchednHQP : public OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> {
public:
    typedef std::shared_ptr<BatchednHQP> Ptr;
    typedef OpenSoT::Solver<Eigen::MatrixXd, Eigen::VectorXd> Base;
    typedef Base::Stack Stack;
    typedef Base::ConstraintPtr ConstraintPtr;
    typedef Base::TaskPtr TaskPtr;
    typedef OpenSoT::solvers::solver_back_ends solver_back_ends;
    typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> RowMat;
    static constexpr double DEFAULT_MIN_SV_RATIO = 0.05;       /* nHQP.h:45 */

    /* the batch: one Stack and one bounds object per instance */
    BatchednHQP(std::vector<Stack>& stacks, const std::vector<ConstraintPtr>& bounds, const double eps_regularisation,
                const int device = 0)
        : Base(stacks.at(0), bounds.at(0))
    {
        if (stacks.size() != bounds.size()) throw std::runtime_error("[nHQP] one bounds object per stack is needed");
        for (size_t i = 0; i < stacks.size(); ++i) { Instance in; in.tasks = stacks[i]; in.bounds = bounds[i]; _inst.push_back(in); }
        init(eps_regularisation, device);
    }
    /* the reference's signature (nHQP.h:56-59): a batch of one */
    BatchednHQP(Stack& stack_of_tasks, ConstraintPtr bounds, const double eps_regularisation,
                const solver_back_ends = solver_back_ends::qpOASES)
        : Base(stack_of_tasks, bounds)
    {
        Instance in; in.tasks = stack_of_tasks; in.bounds = bounds; _inst.push_back(in);
        init(eps_regularisation, 0);
    }
    ~BatchednHQP()
    {
        for (auto& l : _layer) { osb_engine_destroy(l.engine); osb_host_free(l.blobs); osb_host_free(l.z); }
        osb_host_free(_status_pl);
    }
    BatchednHQP(const BatchednHQP&) = delete;
    BatchednHQP& operator=(const BatchednHQP&) = delete;

    /* solutions: n x B, column i = joint velocities of instance i */
    bool solve(Eigen::MatrixXd& solutions)
    {
        const int B = (int)_inst.size(), n = _n, nl = (int)_layer.size();
        for (auto& in : _inst) { in.q0.setZero(n); in.N[0].setIdentity(n, n); }
        _t_prepare = _t_engine = 0.0;
        typedef std::chrono::steady_clock clk;
        for (int i = 0; i < nl; ++i) {
            Layer& l = _layer[i];
            std::atomic<int> bad(0);
            const auto t0 = clk::now();
            pool().run(B, [&](int k) { if (!prepare_layer(i, k)) bad.store(1); });
            const auto t1 = clk::now();
            _t_prepare += std::chrono::duration<double>(t1 - t0).count();
            if (bad.load()) return false;
            const int rc = osb_solve_host(l.engine, l.blobs, B, l.z, _status_pl, nullptr);
            _t_engine += std::chrono::duration<double>(clk::now() - t1).count();
            if (rc != OSB_SUCCESS) { _error = osb_last_error(); return false; }
            std::memcpy(_status.data(), _status_pl, sizeof(int32_t) * (size_t)B);
            for (int k = 0; k < B; ++k)
                if (_status[k] != OSB_OK) { _error = "layer " + std::to_string(i) + ": QP of instance " + std::to_string(k) + " failed"; return false; }
            pool().run(B, [&](int k) {
                Instance& in = _inst[k];
                Eigen::Map<const Eigen::VectorXd> z(l.z + (size_t)k * l.f, l.f);
                in.q0.noalias() += in.N[i] * z;                                     /* solK = solK-1 + NK-1 * xK_opt */
                if (i + 1 < nl) in.N[i + 1].noalias() = in.N[i] * in.Vn;
            });
        }
        if (solutions.rows() != n || solutions.cols() != B) solutions.resize(n, B);
        for (int k = 0; k < B; ++k) solutions.col(k) = _inst[k].q0;
        return true;
    }
    bool solve(Eigen::VectorXd& solution) override
    {
        const bool ok = solve(_X);
        if (solution.size() != _X.size()) solution.resize(_X.size());
        std::memcpy(solution.data(), _X.data(), sizeof(double) * (size_t)_X.size());
        return ok;
    }

    /* nHQP.h:70-110 */
    void setMinSingularValueRatio(double sv_min) { setMinSingularValueRatio(std::vector<double>(_layer.size(), sv_min)); }
    void setMinSingularValueRatio(std::vector<double> sv_min)
    {
        if (sv_min.size() != _layer.size()) throw std::invalid_argument("[nHQP::setMinSingularValueRatio] sv_min.size() != # layers");
        for (size_t i = 0; i < sv_min.size(); ++i) {
            if (sv_min[i] < 0.0 || sv_min[i] > 1.0)
                throw std::invalid_argument("[nHQP::TaskData::set_min_sv_ratio] Minimum singular value threshold should respect 0 < s < 1");
            _layer[i].min_sv_ratio = sv_min[i];
        }
    }
    void setPerformAbRegularization(int hierarchy_level, bool on)
    {
        if (hierarchy_level >= (int)_layer.size()) throw std::invalid_argument("hierarchy_level >= # layers");
        _layer[hierarchy_level].ab_reg = on;
    }
    void setPerformAbRegularization(bool on) { for (auto& l : _layer) l.ab_reg = on; }
    void setPerformSelectiveNullSpaceRegularization(int hierarchy_level, bool on)
    {
        if (hierarchy_level >= (int)_layer.size()) throw std::invalid_argument("hierarchy_level >= # layers");
        _layer[hierarchy_level].ns_reg = on;
    }
    void setPerformSelectiveNullSpaceRegularization(bool on) { for (auto& l : _layer) l.ns_reg = on; }

    unsigned int getNumberOfTasks() { return (unsigned int)_layer.size(); }
    unsigned int getBatchSize() const { return (unsigned int)_inst.size(); }
    int getFreeVariables(int layer) const { return _layer.at(layer).f; }
    const std::string& getLastError() const { return _error; }
    const std::vector<int32_t>& getStatus() const { return _status; }
    void setPackThreads(const int k) { _threads = k < 1 ? 1 : k; _pool.reset(); }
    /* of the last solve(): the per-level preparation on the host threads (SVD, damping, H, g, projected rows) and the C-ABI calls */
    double getLastPrepareSeconds() const { return _t_prepare; }
    double getLastEngineSeconds() const { return _t_engine; }

private:
    struct Instance {
        Stack tasks;
        ConstraintPtr bounds;
        Eigen::VectorXd q0;
        std::vector<Eigen::MatrixXd> N;     /* cumulated null spaces, N[0] = I */
        Eigen::MatrixXd Vn;                 /* null space of the level just prepared */
    };
    struct Layer {
        int f = 0, rows = 0, ns_dim = 0;
        double min_sv_ratio = DEFAULT_MIN_SV_RATIO;
        bool ab_reg = true, ns_reg = true;
        osb_engine* engine = nullptr;
        osb_stack_desc desc;
        osb_layout layout;
        double *blobs = nullptr, *z = nullptr;
    };

    WorkerPool& pool()
    {
        if (!_pool) {
            int k = _threads;
            if (k == 0) { k = (int)std::thread::hardware_concurrency(); if (k > 16) k = 16; if (k < 1) k = 1; }
            _pool.reset(new WorkerPool(k - 1));
        }
        return *_pool;
    }

    /* AN, b0, the SVD with the reference's damping, H and g of one instance at one layer (nHQP.cpp:236-280,357-390) */
    struct Cost { Eigen::MatrixXd AN, H; Eigen::VectorXd b0, g; Eigen::MatrixXd Vn; Eigen::VectorXd sv; };
    static void compute_cost(const TaskPtr& task, const Eigen::MatrixXd* N, const Eigen::VectorXd& q0, const Layer& l, Cost& c)
    {
        /* A = I (Postural-like levels, tasks/velocity/Postural.cpp; normally the last one): AN = N has orthonormal columns, so
         * its singular values are all 1 and U S V' = N I I' -- nothing is damped, b0 keeps its part in range(N) (what
         * U U' b0 with br(i) = 0 beyond the singular values leaves, nHQP.cpp:244-269), and the level has no null space left
         * to hand on. No decomposition: the 29 x 14 SVD of the C29 stack was 60 % of the SVD time of a solve. */
        const int nx = (int)task->getA().cols();
        if (l.ns_dim <= 0 && task->getA().rows() == nx && task->getA().isIdentity(0.0)) {
            if (!N) { c.AN.setIdentity(nx, nx); c.b0 = task->getb(); }
            else {
                c.AN = *N;
                c.b0.noalias() = task->getb() - q0;
                if (l.ab_reg) { const Eigen::VectorXd br = N->transpose() * c.b0; c.b0.noalias() = (*N) * br; }
            }
            c.sv.setOnes(c.AN.cols());
            c.H.noalias() = c.AN.transpose() * task->getWeight() * c.AN;
            c.g.noalias() = -c.AN.transpose() * (task->getWeight() * c.b0);     /* (matrix-vector products, not AN' W first) */
            c.Vn.resize(c.AN.cols(), 0);
            return;
        }
        if (!N) { c.AN = task->getA(); c.b0 = task->getb(); }
        else { c.AN.noalias() = task->getA() * (*N); c.b0.noalias() = task->getb() - task->getA() * q0; }
        Eigen::JacobiSVD<Eigen::MatrixXd> svd(c.AN, Eigen::ComputeFullU | Eigen::ComputeFullV);
        c.sv = svd.singularValues();
        const double sv_max = c.sv.size() ? c.sv(0) : 0.0;
        if (l.ab_reg) {
            Eigen::VectorXd sv = c.sv;
            Eigen::VectorXd br = svd.matrixU().transpose() * c.b0;
            const double thr = l.min_sv_ratio;
            for (int i = 0; i < br.size(); ++i) {
                if (i >= sv.size()) br(i) = 0.0;
                else if (sv(i) < thr * sv_max) {
                    br(i) *= sv(i) / (thr * sv_max);
                    sv(i) = std::pow(thr * sv_max, 2u) / (sv(i) + thr / 100.0);
                }
            }
            c.b0 = svd.matrixU() * br;
            c.AN = svd.matrixU().leftCols(sv.size()) * sv.asDiagonal() * svd.matrixV().transpose().topRows(sv.size());
        }
        c.H.noalias() = c.AN.transpose() * task->getWeight() * c.AN;
        c.g.noalias() = -c.AN.transpose() * (task->getWeight() * c.b0);     /* (matrix-vector products, not AN' W first) */
        if (l.ns_dim > 0) {
            c.Vn = svd.matrixV().rightCols(l.ns_dim);
            if (l.ns_reg) c.H.noalias() += sv_max * c.Vn * c.Vn.transpose();
        } else c.Vn.resize(c.AN.cols(), 0);
    }

    void init(const double eps_regularisation, const int device)
    {
        Instance& i0 = _inst[0];
        if (i0.tasks.empty()) throw std::runtime_error("[nHQP] empty stack");
        _n = (int)i0.tasks.front()->getXSize();
        const int n = _n, nl = (int)i0.tasks.size(), B = (int)_inst.size();
        for (auto& in : _inst) {
            if ((int)in.tasks.size() != nl) throw std::runtime_error("[nHQP] stacks differ in shape");
            if (in.bounds->getAeq().rows() > 0) throw std::runtime_error("[nHQP] Equality constraints not supported");
            in.N.assign((size_t)nl, Eigen::MatrixXd());
            in.N[0].setIdentity(n, n);
            in.q0.setZero(n);
        }
        const int c_rows = (int)i0.bounds->getAineq().rows();
        int f = n;
        _layer.resize((size_t)nl);
        for (int i = 0; i < nl; ++i) {
            for (auto& in : _inst) {
                in.tasks[i]->update();                                              /* nHQP.cpp:30 */
                if (in.tasks[i]->getConstraints().size() > 0) throw std::runtime_error("[nHQP] Local constraints not supported");
            }
            if (f <= 0) throw std::runtime_error("[nHQP] No free variables left at layer #" + std::to_string(i) + ": decrease the number of layers!");
            Layer& l = _layer[i];
            l.f = f;
            l.rows = c_rows + (i == 0 ? 0 : n);                                     /* bounds become rows below the first layer */
            l.ns_dim = f - (int)i0.tasks[i]->getTaskSize();
            std::memset(&l.desc, 0, sizeof(l.desc));
            l.desc.n = f; l.desc.n_levels = 1; l.desc.m[0] = 0; l.desc.task_kind[0] = OSB_TASK_HESSIAN;
            l.desc.c_in = l.rows; l.desc.n_bound_sets = (i == 0 && i0.bounds->hasBounds()) ? 1 : 0;
            l.desc.eps = eps_regularisation * OSB_QPOASES_BASE_EPS;
            if (osb_blob_layout(&l.desc, &l.layout) != OSB_SUCCESS) throw std::runtime_error("[nHQP] layer shape not supported by the engine");
            if (osb_engine_create(&l.engine, &l.desc, B, device) != OSB_SUCCESS) throw std::runtime_error(std::string("[nHQP] ") + osb_last_error());
            l.blobs = (double*)osb_host_alloc((int64_t)sizeof(double) * B * l.layout.stride);
            l.z = (double*)osb_host_alloc((int64_t)sizeof(double) * B * f);
            if (!l.blobs || !l.z) throw std::runtime_error("[nHQP] page-locked host memory could not be allocated");
            std::memset(l.blobs, 0, sizeof(double) * (size_t)B * l.layout.stride);
            if (i == nl - 1) continue;
            /* null-space dimension of the layer from the rank of the first instance, no regularisation
             * (nHQP.cpp:81-113, SV_THRESH = 1e-6); every instance must agree: the batch shares its shapes */
            int ns = -1;
            for (int k = 0; k < B; ++k) {
                Instance& in = _inst[k];
                Cost c; Layer probe = l; probe.ab_reg = false; probe.ns_reg = false; probe.ns_dim = 0;
                compute_cost(in.tasks[i], i == 0 ? nullptr : &in.N[i], in.q0, probe, c);
                const int rank = (int)(c.sv.array() >= 1e-6).count();
                const int nsk = f - rank;
                if (ns < 0) ns = nsk;
                else if (nsk != ns) throw std::runtime_error("[nHQP] instances differ in the rank of layer #" + std::to_string(i));
                Eigen::JacobiSVD<Eigen::MatrixXd> svd(c.AN, Eigen::ComputeFullV);
                if (nsk <= 0) throw std::runtime_error("Nullspace basis not available");
                in.N[i + 1] = in.N[i] * svd.matrixV().rightCols(nsk);
            }
            l.ns_dim = ns;
            f = ns;
        }
        _status.assign((size_t)B, 0);
        osb_host_free(_status_pl);
        _status_pl = (int32_t*)osb_host_alloc((int64_t)sizeof(int32_t) * B);   /* (pageable status words serialise the pipeline of osb_solve_host) */
    }

    /* blob of instance k at layer i: H, g, the projected rows and (first layer) the box */
    bool prepare_layer(const int i, const int k)
    {
        Layer& l = _layer[i];
        Instance& in = _inst[k];
        const int n = _n, f = l.f;
        Cost c;
        compute_cost(in.tasks[i], i == 0 ? nullptr : &in.N[i], in.q0, l, c);
        in.Vn = c.Vn;
        double* blob = l.blobs + (size_t)k * l.layout.stride;
        Eigen::Map<RowMat>(blob + l.layout.A[0], f, f) = c.H;
        Eigen::Map<Eigen::VectorXd>(blob + l.layout.b[0], f) = c.g;
        ConstraintType& con = *in.bounds;
        const int cr = (int)con.getAineq().rows();
        if (cr + (i == 0 ? 0 : n) != l.rows) return false;
        const double inf = std::numeric_limits<double>::infinity();
        auto lo_of = [&](int r) { return con.getbLowerBound().size() ? con.getbLowerBound()[r] : -inf; };
        auto hi_of = [&](int r) { return con.getbUpperBound().size() ? con.getbUpperBound()[r] : inf; };
        if (i == 0) {                                                               /* nHQP.cpp:297-306 */
            if (cr > 0) Eigen::Map<RowMat>(blob + l.layout.Cin, cr, f) = con.getAineq();
            for (int r = 0; r < cr; ++r) { blob[l.layout.lA + r] = lo_of(r); blob[l.layout.uA + r] = hi_of(r); }
            if (l.desc.n_bound_sets) {
                Eigen::Map<Eigen::VectorXd>(blob + l.layout.l[0], n) = con.getLowerBound();
                Eigen::Map<Eigen::VectorXd>(blob + l.layout.u[0], n) = con.getUpperBound();
            }
        } else {                                                                    /* nHQP.cpp:307-317 */
            const Eigen::MatrixXd& N = in.N[i];
            if (cr > 0) {
                Eigen::Map<RowMat>(blob + l.layout.Cin, cr, f) = con.getAineq() * N;
                const Eigen::VectorXd cq = con.getAineq() * in.q0;
                for (int r = 0; r < cr; ++r) { blob[l.layout.lA + r] = lo_of(r) - cq[r]; blob[l.layout.uA + r] = hi_of(r) - cq[r]; }
            }
            Eigen::Map<RowMat>(blob + l.layout.Cin + (size_t)cr * f, n, f) = N;
            for (int r = 0; r < n; ++r) {
                blob[l.layout.lA + cr + r] = (con.getLowerBound().size() ? con.getLowerBound()[r] : -inf) - in.q0[r];
                blob[l.layout.uA + cr + r] = (con.getUpperBound().size() ? con.getUpperBound()[r] : inf) - in.q0[r];
            }
        }
        return true;
    }

    typedef Base::ConstraintType ConstraintType;
    std::vector<Instance> _inst;
    std::vector<Layer> _layer;
    int _n = 0, _threads = 0;
    double _t_prepare = 0.0, _t_engine = 0.0;
    std::vector<int32_t> _status;
    int32_t* _status_pl = nullptr;
    std::unique_ptr<WorkerPool> _pool;
    Eigen::MatrixXd _X;
    std::string _error;
};

}}

#endif

/*
 * opensot_b200.h -- C ABI of the B200-native batched Stack-of-Tasks QP engine.
 *
 * This is the drop-in boundary of the hot path: everything a host
 * (the C++/Eigen front-end in opensot_b200/host, the BackEnd plugin, ctypes, cgo ...) needs
 * to run  "stack A,b  ->  H = A'WA, g = -A'Wb per level  ->  hierarchical QP cascade"
 * for a batch of B structurally identical stacks on one B200.
 * Plain pointers and sizes only; no C++/torch/Eigen types cross this line.
 *
 * Reference interfaces replaced (paths relative to the OpenSoT tree):
 *   - osb_stack_desc / osb_blob_layout : the frozen problem shape that iHQP::prepareSoT
 *       derives once per solver object               src/solvers/iHQP.cpp:172-261
 *       and the row-stacking done by MatrixPiler     include/OpenSoT/utils/Piler.h:102-163
 *       tasks::Aggregated::generateAll               src/tasks/Aggregated.cpp:113-132
 *       constraints::Aggregated::generateAll         src/constraints/Aggregated.cpp:73-257
 *   - osb_engine_create                : BackEndFactory(...) + BackEnd::initProblem per level
 *                                                    src/solvers/BackEndFactory.cpp:19-89
 *   - osb_solve_host / osb_solve_device: iHQP::solve (level loop, computeCostFunction,
 *       computeOptimalityConstraint, BackEnd::updateTask/updateConstraints/updateBounds/
 *       solve/getSolution)                           src/solvers/iHQP.cpp:129-170,263-358
 *                                                    src/solvers/QPOasesBackEnd.cpp:248-307
 *   - osb_set_active_levels            : iHQP::setActiveStack / activateAllStacks
 *                                                    src/solvers/iHQP.cpp:391-400
 *   - osb_stack_gather_device          : AutoStack::update() piling of child tasks / bounds
 *                                                    src/utils/AutoStack.cpp:385-393
 *   - status words                     : bool returns of BackEnd::solve / iHQP::solve
 *                                                    src/solvers/iHQP.cpp:279,335-347
 *   - osb_produce_posture_device / osb_produce_kinematics_device / osb_solve_states_host : the per-cycle update() of
 *       Postural, JointLimits, VelocityLimits, Cartesian, CoM   src/tasks/velocity/Postural.cpp:97-100, Cartesian.cpp:68-105,
 *                                                    CoM.cpp:59-74, src/constraints/velocity/JointLimits.cpp:37-59, VelocityLimits.cpp:72-89
 *   - osb_solve_host_submit / osb_solve_states_submit / osb_solve_wait, osb_wire_configure / osb_solve_host_wire :
 *       no counterpart (the reference solves synchronously and re-piles constant fields every update()); transfer-side
 *       additions of this boundary: two calls in flight, and fields that never change uploaded once
 *
 * Per-instance input ("blob"): one contiguous run of doubles per instance, instance-major,
 * 16-byte aligned, fetched by one cp.async.bulk per level by the solving warp. Field order
 * and offsets come from osb_blob_layout(). Matrices are ROW-major (a level made of several
 * child tasks is therefore just their row blocks one after another: piling is a layout, not
 * a copy).
 */
#ifndef OPENSOT_B200_H_
#define OPENSOT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSB_MAX_LEVELS      8
#define OSB_MAX_BOUND_SETS  4
#define OSB_MAX_PEERS       8   /* GPUs of one NVSwitch domain a fused all-gather spans   */
#define OSB_IPC_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t)                             */
#define OSB_MAX_N           64      /* variables per instance                          */
#define OSB_MAX_ROWS        192     /* n + inequality rows + task rows of one level    */
#define OSB_INFTY           1.0e20  /* qpOASES INFTY; QPOasesBackEnd.cpp:339-356       */
#define OSB_QPOASES_BASE_EPS 2.221e-13 /* 1e3*EPS with qpOASES' EPS = 2.221e-16 (Constants.hpp:48-50; Options.cpp:143-148) */
#define OSB_F32_MIN_EPS     1.0e-6  /* smallest diagonal shift an fp32 engine uses     */

/* how the cost of a level is given */
enum osb_task_kind {
    OSB_TASK_DENSE    = 0,  /* A (m x n row-major), b (m)                              */
    OSB_TASK_IDENTITY = 1,  /* A = I_n implied (Postural-like), b (n); m must equal n  */
    OSB_TASK_HESSIAN  = 2   /* H (n x n), g (n) given directly (BackEnd::updateTask);  */
                            /* only legal on the last level (no A for optimality rows) */
};
enum osb_weight_kind {
    OSB_W_IDENTITY = 0,
    OSB_W_DIAG     = 1,     /* w (m)                                                   */
    OSB_W_DENSE    = 2      /* W (m x m), symmetric positive definite                  */
};

/* per-instance status word (low byte = code, bits 8..15 = level at which it happened) */
enum osb_status {
    OSB_OK            = 0,
    OSB_MAX_ITER      = 1,  /* working-set changes exceeded osb_options.max_iter       */
    OSB_INFEASIBLE    = 2,  /* constraints of some level are inconsistent              */
    OSB_NOT_SPD       = 3,  /* reduced Hessian not positive definite (eps == 0 ?)      */
    OSB_NAN_INPUT     = 4
};

/* error codes returned by the API functions (0 = success) */
enum osb_error {
    OSB_SUCCESS = 0, OSB_E_INVALID = -1, OSB_E_CUDA = -2, OSB_E_TOO_LARGE = -3,
    OSB_E_NO_DEVICE = -4,
    OSB_E_PRECISION = -5    /* fp32 engine asked for a stack whose level Hessians single precision cannot hold */
};

/* Shape of one stack; identical for every instance of a batch. */
typedef struct osb_stack_desc {
    int32_t n;                              /* variables                               */
    int32_t n_levels;                       /* priority levels, 1..OSB_MAX_LEVELS      */
    int32_t m[OSB_MAX_LEVELS];              /* task rows of level k                    */
    int32_t task_kind[OSB_MAX_LEVELS];      /* osb_task_kind                           */
    int32_t weight_kind[OSB_MAX_LEVELS];    /* osb_weight_kind                         */
    int32_t has_c[OSB_MAX_LEVELS];          /* linear cost term c (n), Task::getc()    */
    int32_t c_loc[OSB_MAX_LEVELS];          /* task-local two-sided inequality rows    */
    int32_t c_eq;                           /* global equality rows   Ceq x = deq      */
    int32_t c_in;                           /* global two-sided rows  lA <= Cin x <= uA*/
    int32_t n_bound_sets;                   /* (l,u) pairs, intersected on device      */
    int32_t m_reg;                          /* rows of the user regularisation task    */
    int32_t reg_weight_kind;                /* OSB_W_IDENTITY or OSB_W_DIAG            */
    double  eps;                            /* ABSOLUTE diagonal shift added to every  */
                                            /* level Hessian (factor * 2.221e-13)       */
} osb_stack_desc;

/* Offsets (in doubles) of every field inside one instance blob; -1 = field absent. */
typedef struct osb_layout {
    int32_t A[OSB_MAX_LEVELS];      /* DENSE: m*n ; HESSIAN: n*n ; IDENTITY: -1        */
    int32_t b[OSB_MAX_LEVELS];      /* m (HESSIAN: g, n)                               */
    int32_t w[OSB_MAX_LEVELS];      /* DIAG: m ; DENSE: m*m                            */
    int32_t c[OSB_MAX_LEVELS];      /* n                                               */
    int32_t Cl[OSB_MAX_LEVELS];     /* c_loc*n                                         */
    int32_t lCl[OSB_MAX_LEVELS];    /* c_loc                                           */
    int32_t uCl[OSB_MAX_LEVELS];    /* c_loc                                           */
    int32_t level_begin[OSB_MAX_LEVELS]; /* first double of level k's block            */
    int32_t level_size[OSB_MAX_LEVELS];  /* doubles in level k's block (even)          */
    int32_t Ceq, deq;               /* c_eq*n, c_eq                                    */
    int32_t Cin, lA, uA;            /* c_in*n, c_in, c_in                              */
    int32_t l[OSB_MAX_BOUND_SETS];  /* n each                                          */
    int32_t u[OSB_MAX_BOUND_SETS];
    int32_t Ar, br, wr;             /* regularisation task: m_reg*n, m_reg, m_reg      */
    int32_t global_begin;           /* first double of the global block                */
    int32_t global_size;            /* doubles in the global block (even)              */
    int32_t stride;                 /* doubles per instance (even => 16 B aligned)     */
} osb_layout;

typedef struct osb_options {
    int32_t max_iter;       /* working-set changes per level; <=0 -> 10*(n+rows)       */
    int32_t store_levels;   /* also return the solution of every level                 */
    double  tol_feas;       /* constraint violation tolerance (absolute, scaled by     */
                            /* max(1,|bound|)); <=0 -> 1e-12                           */
    double  tol_rank;       /* relative row-dependence tolerance; <=0 -> 1e-11         */
} osb_options;

typedef struct osb_engine osb_engine;

/* Validates the descriptor and fills the layout. Pure host function, no CUDA. */
int  osb_blob_layout(const osb_stack_desc* desc, osb_layout* out);
/* Shared memory (bytes) one solving warp needs for this shape; pure host function. */
int64_t osb_smem_bytes_per_instance(const osb_stack_desc* desc);

/* sizeof of {osb_stack_desc, osb_layout, osb_options, osb_engine_info, osb_gather_src,
 * osb_posture_inputs}: lets
 * a foreign-language binding verify its struct mirrors before the first real call. */
void osb_struct_sizes(int32_t out[6]);
const char* osb_last_error(void);
const char* osb_version(void);
int  osb_device_count(void);

int  osb_engine_create(osb_engine** out, const osb_stack_desc* desc, int32_t max_batch,
                       int32_t device);
/* Single-precision engine (BASELINE cfg "MPC-style ... fp32"): same stack semantics, blobs,
 * solutions and all arithmetic in float; instance blobs are osb_blob_stride_f32() floats apart
 * (field offsets unchanged, stride rounded up to a multiple of 4 for 16-byte alignment).
 * Default tolerances become tol_feas 1e-6, tol_rank 1e-5, and desc->eps is raised to
 * OSB_F32_MIN_EPS if smaller. The reference has no fp32 path
 * (QPOasesBackEnd hands qpOASES double data, QPOasesBackEnd.cpp:129-133): parity of this mode
 * is judged against the fp64 oracle with the tolerance stated in tests/test_gpu_parity.py. */
int  osb_engine_create_f32(osb_engine** out, const osb_stack_desc* desc, int32_t max_batch,
                           int32_t device);
int32_t osb_blob_stride_f32(const osb_stack_desc* desc);
void osb_engine_destroy(osb_engine* e);
int  osb_set_options(osb_engine* e, const osb_options* opt);
/* active[k] != 0 <=> level k takes part; inactive levels contribute neither a QP nor
 * optimality rows (iHQP.cpp:296-310). */
int  osb_set_active_levels(osb_engine* e, const int32_t* active);

/* Host-buffer entry (what the reference-facing plugin calls every control cycle):
 * blobs[batch*stride] -> x[batch*n], status[batch]. Copies H2D, solves, copies D2H,
 * synchronises. x_levels may be NULL (else [batch*n_levels*n], needs store_levels). */
int  osb_solve_host(osb_engine* e, const double* blobs, int32_t batch,
                    double* x, int32_t* status, double* x_levels);

/* Device-resident entry: all pointers are device pointers on the engine's device; the
 * work is enqueued on `cuda_stream` (a cudaStream_t, NULL = legacy default stream) and
 * NOT synchronised. iters (int32[batch], total working-set changes) may be NULL. */
int  osb_solve_device(osb_engine* e, const double* d_blobs, int32_t batch,
                      double* d_x, int32_t* d_status, double* d_x_levels,
                      int32_t* d_iters, void* cuda_stream);

/* fp32 twins of the two solve entry points (engine from osb_engine_create_f32) */
int  osb_solve_host_f32(osb_engine* e, const float* blobs, int32_t batch,
                        float* x, int32_t* status, float* x_levels);
int  osb_solve_device_f32(osb_engine* e, const float* d_blobs, int32_t batch,
                          float* d_x, int32_t* d_status, float* d_x_levels,
                          int32_t* d_iters, void* cuda_stream);

/* Two host-buffer calls in flight (throughput mode for callers whose batches do not depend on each other: MPC samples,
 * roll-outs, or a control loop that accepts the solution one cycle later). osb_solve_host_submit enqueues what
 * osb_solve_host does and returns without waiting; osb_solve_wait blocks until the OLDEST submitted call has delivered
 * x / status (/ x_levels) to its host buffers. At most two calls may be in flight (a third submit fails with
 * OSB_E_INVALID); each uses its own set of device buffers, so the blobs of call k+1 cross PCIe while call k is still
 * being solved and its solutions travel back -- the step time becomes the PCIe transfer alone instead of transfer +
 * last solve + copy-out. Host buffers must stay valid (page-locked: osb_host_alloc) and untouched until the matching
 * wait; the blocking entry points refuse to run while submitted calls are in flight. Unlike osb_solve_host a submitted
 * call is NOT ordered behind work the caller enqueued on osb_engine_stream().
 * The reference has no counterpart: iHQP::solve is synchronous (iHQP.cpp:263-358). */
int  osb_solve_host_submit(osb_engine* e, const double* blobs, int32_t batch,
                           double* x, int32_t* status, double* x_levels);
int  osb_solve_host_submit_f32(osb_engine* e, const float* blobs, int32_t batch,
                               float* x, int32_t* status, float* x_levels);
int  osb_solve_wait(osb_engine* e);
int32_t osb_solve_in_flight(osb_engine* e);     /* 0, 1 or 2 */

/* Wire format: only what changes every cycle crosses PCIe. A stack usually has fields that never change between cycles
 * (weights, velocity limits, constant constraint matrices: in the reference they are set once on the task objects and
 * re-piled by every update(), Aggregated.cpp:113-132). osb_wire_configure declares them: `static_segments` are sorted,
 * disjoint [offset, offset + count) ranges of a blob (in elements, osb_layout offsets), `template_blob` (layout.stride
 * doubles) carries their values. A wire row is the remaining (dynamic) runs of the blob, packed in order, osb_wire_stride()
 * elements of the engine's precision long (rounded up to 16 bytes); osb_wire_pack builds wire rows from full blobs on the
 * host. osb_solve_host_wire / osb_solve_host_wire_submit are osb_solve_host / osb_solve_host_submit for wire rows: H2D of
 * the wire rows, one scatter kernel into the blobs (whose static part was written once), solve, D2H. Element type of
 * wire / x / x_levels = the engine's precision. */
typedef struct osb_wire_segment { int32_t offset, count; } osb_wire_segment;
int  osb_wire_configure(osb_engine* e, const osb_wire_segment* static_segments, int32_t n_static, const double* template_blob);
int32_t osb_wire_stride(osb_engine* e);
int  osb_wire_pack(osb_engine* e, const void* blobs, int32_t batch, void* wire);
int  osb_solve_host_wire(osb_engine* e, const void* wire, int32_t batch, void* x, int32_t* status, void* x_levels);
int  osb_solve_host_wire_submit(osb_engine* e, const void* wire, int32_t batch, void* x, int32_t* status, void* x_levels);

/* Aggregated/AutoStack::update() on device: gathers per-task row blocks and bound sets
 * living in separate [B, ...] device arrays into instance blobs.
 * For every source s: rows[s] x cols[s] doubles per instance at src[s] + i*src_stride[s],
 * copied to blob offset dst_off[s] of instance i, multiplied by scale[s]
 * (SubTask lambda / task weight folded in; SubTask.cpp:57-58). row_index[s] (device int32
 * array of rows[s] entries, or NULL) selects source rows (operator %). */
typedef struct osb_gather_src {
    const double*  src;         /* device                                              */
    int64_t        src_stride;  /* doubles between consecutive instances               */
    int32_t        rows, cols;  /* block copied per instance (rows*cols doubles)       */
    int32_t        src_ld;      /* leading dimension (doubles) of a source row         */
    const int32_t* row_index;   /* device, may be NULL                                 */
    int32_t        dst_off;     /* offset in the blob (doubles)                        */
    double         scale;
} osb_gather_src;
int  osb_stack_gather_device(osb_engine* e, const osb_gather_src* srcs, int32_t n_srcs,
                             double* d_blobs, int32_t batch, void* cuda_stream);

/* On-device producers of the inputs that depend on the joint state only (SURVEY.md section 8f,
 * rank 1), so that a batch of robot STATES, not of pre-computed vectors, feeds the blobs:
 *   Postural::update_b        b = v_ref + lambda (q_ref - q)          src/tasks/velocity/Postural.cpp:97-100
 *   JointLimits::update       u = max(0, (q_max - (q - q_neutral)) s)  src/constraints/velocity/JointLimits.cpp:37-59
 *                             l = min(0, (q_min - (q - q_neutral)) s)
 *   VelocityLimits::generateBounds   u = |qdot_max| dT, l = -u         src/constraints/velocity/VelocityLimits.cpp:72-89
 * for Euclidean joint spaces (nq == nv: ModelInterface::difference is a subtraction). Every
 * pointer is a device pointer; q and q_ref are [batch, n] with the given strides in doubles
 * (stride 0 = one vector shared by the batch), v_ref may be NULL (zero), the limit vectors are
 * [n]; q_neutral may be NULL (zero). postural_level < 0, joint_limit_set < 0 or
 * velocity_limit_set < 0 skip that producer. */
typedef struct osb_posture_inputs {
    const double* q;        int64_t q_stride;
    const double* q_ref;    int64_t q_ref_stride;
    const double* v_ref;    int64_t v_ref_stride;
    const double* q_min;    const double* q_max;    const double* q_neutral;
    const double* qdot_max;
    double  lambda, bound_scaling, dT;
    int32_t postural_level;        /* an OSB_TASK_IDENTITY level: receives b                     */
    int32_t joint_limit_set;       /* bound set receiving the JointLimits box                    */
    int32_t velocity_limit_set;    /* bound set receiving the VelocityLimits box                 */
    int32_t reserved;
} osb_posture_inputs;
int  osb_produce_posture_device(osb_engine* e, const osb_posture_inputs* in, double* d_blobs,
                                int32_t batch, void* cuda_stream);

/*
 * Cartesian and CoM producers (rest): A and b of tasks::velocity::Cartesian
 * (src/tasks/velocity/Cartesian.cpp:68-105, 279-285) and tasks::velocity::CoM (src/tasks/velocity/CoM.cpp:59-74,
 * 145-149) written into the blobs from a batch of joint positions and a kinematic tree, so that together with
 * osb_produce_posture_device only joint states cross PCIe. The tree is what a URDF describes (the reference gets
 * it through XBot::ModelInterface): every joint, fixed ones included, parents before children; joint j carries the
 * frame of its child link. Fixed-base trees of revolute / prismatic / fixed joints.
 */
enum osb_kin_joint_type { OSB_KIN_FIXED = 0, OSB_KIN_REVOLUTE = 1, OSB_KIN_PRISMATIC = 2 };
enum osb_kin_task_kind { OSB_KIN_CARTESIAN = 0, OSB_KIN_COM = 1 };
typedef struct osb_kin_tree {       /* host arrays, copied to the device by osb_kinematics_create */
    int32_t n_joints, n_dof;
    const int32_t* parent;          /* [n_joints] index of the parent joint, -1 = attached to the world      */
    const int32_t* type;            /* [n_joints] osb_kin_joint_type                                         */
    const int32_t* dof;             /* [n_joints] column of q / of the Jacobian, -1 for fixed joints         */
    const double*  axis;            /* [n_joints][3] joint axis in the joint frame                           */
    const double*  origin;          /* [n_joints][12] parent link -> joint frame: rotation (row-major), xyz  */
    const double*  mass;            /* [n_joints] mass of the child link                                     */
    const double*  com;             /* [n_joints][3] its centre of mass in the link frame                    */
} osb_kin_tree;
typedef struct osb_kin_task_desc {
    int32_t kind;                   /* osb_kin_task_kind                                                     */
    int32_t distal, base;           /* joints whose child links are the distal / base link; base < 0: world  */
    int32_t level, row0;            /* destination: rows [row0, row0 + 6 or 3) of the OSB_TASK_DENSE level   */
    int32_t reserved;
    double  lambda, orientation_gain;
    const double* desired;          /* device, [batch] x (pose: rotation row-major 9 + position 3 | CoM: 3); */
    int64_t desired_stride;         /*   stride in doubles, 0 = shared; NULL = zero error                    */
    const double* desired_vel;      /* device, [batch] x (6 | 3) feed-forward twist / velocity, or NULL      */
    int64_t desired_vel_stride;
} osb_kin_task_desc;
int  osb_kinematics_create(osb_engine* e, const osb_kin_tree* tree);
int  osb_produce_kinematics_device(osb_engine* e, const double* d_q, int64_t q_stride, const osb_kin_task_desc* tasks,
                                   int32_t n_tasks, double* d_blobs, int32_t batch, void* cuda_stream);

/*
 * State-input path: the whole per-cycle update of a stack made of Postural / JointLimits / VelocityLimits /
 * Cartesian / CoM pieces runs on the device, so a control cycle uploads joint positions only
 * (batch x n doubles instead of batch x stride). Configure once with HOST data (copied to the device and kept):
 * a blob template holding every field that does not depend on q (weights, constant rows), the limit vectors and
 * references, and the kinematic tasks (their `desired` / `desired_vel` are host arrays shared by the batch here).
 * osb_solve_states_host then is: H2D q -> sot_posture_kernel -> sot_kinematics_kernel -> solver kernel -> D2H x.
 */
typedef struct osb_states_config {
    const double* blob_template;    /* host, layout.stride doubles                                           */
    const double* q_ref; const double* q_min; const double* q_max; const double* q_neutral; const double* qdot_max; /* host [n] or NULL */
    double  lambda, bound_scaling, dT;
    int32_t postural_level, joint_limit_set, velocity_limit_set;   /* < 0: producer skipped                  */
    int32_t n_tasks;
    const osb_kin_task_desc* tasks; /* host; desired: 12 (pose) or 3 (CoM) doubles, desired_vel: 6 or 3       */
} osb_states_config;
int  osb_states_configure(osb_engine* e, const osb_states_config* cfg);
int  osb_solve_states_host(osb_engine* e, const double* q, int32_t batch, double* x, int32_t* status);
/* submitted twin (see osb_solve_host_submit; completed by osb_solve_wait) */
int  osb_solve_states_submit(osb_engine* e, const double* q, int32_t batch, double* x, int32_t* status);
/* the blobs the last osb_solve_states_host built on the device, copied back (debugging / parity tests) */
int  osb_states_download_blobs(osb_engine* e, double* blobs, int32_t batch);

/* Introspection for tests / bench. */
typedef struct osb_engine_info {
    int32_t device, sm_count, warps_per_cta, ctas_per_sm, grid;
    int64_t smem_per_cta, blob_stride_bytes;
    int32_t kernel_regs;
    int64_t launches;           /* kernels launched by this engine so far              */
} osb_engine_info;
int  osb_engine_get_info(osb_engine* e, osb_engine_info* out);
int  osb_engine_device_buffers(osb_engine* e, double** d_blobs, double** d_x,
                               int32_t** d_status);
/* the cudaStream_t osb_solve_host works on (so that a caller can record events on it) */
void* osb_engine_stream(osb_engine* e);
/* which solver kernel this engine launches: "rr_solve_kernel<f64,30x1>" (register rows, 30
 * columns, one row per lane) ... or "sot_solve_kernel" (generic shapes); see DESIGN.md section 4 */
const char* osb_engine_kernel_name(osb_engine* e);

/*
 * Solve fused with the all-gather of the solutions, over peer memory (NVLink 5 / NVSwitch).
 * The batch shards over the GPUs of one box, one process per GPU; when the
 * caller needs the whole batch on every rank, the solve kernel itself stores each x row and
 * status word into the gather buffer of every rank as the stack finishes (plain peer stores,
 * posted over NVLink while the other stacks are still being solved), instead of an NCCL
 * all-gather after the kernel. Nothing in the reference to cite: it solves one robot per process.
 *   osb_gather_create   allocates this rank's gather buffer: x_all [global_batch, n] in the
 *                       engine's precision, status_all [global_batch] int32, a flag block;
 *                       writes its CUDA IPC handle (OSB_IPC_HANDLE_BYTES) for the other ranks.
 *                       This rank solves global instances [first_instance, first_instance + batch).
 *   osb_gather_connect  opens the buffer of rank `peer` from its handle (exchange the handles
 *                       with any host-side all-gather, e.g. torch.distributed).
 *   osb_solve_device_gather  one launch of the solver kernel with the peer stores enabled.
 *   osb_gather_barrier  stream-ordered device-side barrier over the flag blocks: when it has
 *                       run, the rows of EVERY rank launched before their barrier are visible in
 *                       this rank's buffer. Fails (status word, no hang) after timeout_ms.
 *   osb_gather_buffers  this rank's x_all / status_all device pointers. The buffer has TWO halves used by step
 *                       parity: the pointers returned are those of the LAST launched fused solve, the next one
 *                       gathers into the other half, so a rank still reading step s is never overwritten by a peer
 *                       that has already started step s + 1 (call it again after every osb_solve_device_gather).
 */
int  osb_gather_create(osb_engine* e, int64_t global_batch, int32_t rank, int32_t world,
                       int64_t first_instance, uint8_t handle_out[OSB_IPC_HANDLE_BYTES]);
int  osb_gather_connect(osb_engine* e, int32_t peer, const uint8_t handle[OSB_IPC_HANDLE_BYTES]);
int  osb_solve_device_gather(osb_engine* e, const void* d_blobs, int32_t batch, void* cuda_stream);
int  osb_gather_barrier(osb_engine* e, void* cuda_stream, int32_t timeout_ms);
int  osb_gather_buffers(osb_engine* e, void** d_x_all, int32_t** d_status_all);
/* 0 while every barrier so far completed, else the number of barriers that timed out */
int  osb_gather_barrier_failures(osb_engine* e);

/* Absolute diagonal shift of one level (level < 0: every level). Replaces BackEnd::setEpsRegularisation as
 * driven by iHQP::setEpsRegularisation(eps[, i]) (src/solvers/iHQP.cpp:437-459, QPOasesBackEnd.cpp:358-372).
 * fp32 engines keep at least OSB_F32_MIN_EPS. */
int    osb_set_eps_level(osb_engine* e, int32_t level, double eps);
double osb_get_eps_level(osb_engine* e, int32_t level);
/* the options in force (BackEnd::getOptions, src/solvers/QPOasesBackEnd.cpp) */
int    osb_get_options(osb_engine* e, osb_options* out);
/* page-locked host buffers for osb_solve_host (blobs in, solutions out); NULL on failure */
void*  osb_host_alloc(int64_t bytes);
void   osb_host_free(void* ptr);

/* FP64 FMA peak probe used for the roofline denominator (bench.py): runs a register-only
 * DFMA kernel and returns achieved TFLOP/s (2 flop per FMA), <0 on error. */
double osb_measure_fp64_tflops(int32_t device, int32_t iters);
/* the same with FFMA, for the fp32 engines */
double osb_measure_fp32_tflops(int32_t device, int32_t iters);

#ifdef __cplusplus
}
#endif
#endif /* OPENSOT_B200_H_ */

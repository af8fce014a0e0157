"""ctypes mirror of include/opensot_b200.h (structs, enums, constants). No CUDA, no torch."""
import ctypes as C

OSB_MAX_LEVELS = 8
OSB_MAX_BOUND_SETS = 4
OSB_MAX_PEERS = 8
OSB_IPC_HANDLE_BYTES = 64
OSB_MAX_N = 64
OSB_MAX_ROWS = 192
OSB_INFTY = 1.0e20
OSB_QPOASES_BASE_EPS = 2.221e-13
OSB_F32_MIN_EPS = 1.0e-6
F32_EPS_FACTOR = OSB_F32_MIN_EPS / OSB_QPOASES_BASE_EPS

TASK_DENSE, TASK_IDENTITY, TASK_HESSIAN = 0, 1, 2
W_IDENTITY, W_DIAG, W_DENSE = 0, 1, 2
OK, MAX_ITER, INFEASIBLE, NOT_SPD, NAN_INPUT = 0, 1, 2, 3, 4

_I8 = C.c_int32 * OSB_MAX_LEVELS
_I4 = C.c_int32 * OSB_MAX_BOUND_SETS


class StackDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_levels", C.c_int32),
        ("m", _I8), ("task_kind", _I8), ("weight_kind", _I8), ("has_c", _I8), ("c_loc", _I8),
        ("c_eq", C.c_int32), ("c_in", C.c_int32), ("n_bound_sets", C.c_int32),
        ("m_reg", C.c_int32), ("reg_weight_kind", C.c_int32),
        ("eps", C.c_double),
    ]

    def key(self):
        return bytes(self)


class Layout(C.Structure):
    _fields_ = [
        ("A", _I8), ("b", _I8), ("w", _I8), ("c", _I8), ("Cl", _I8), ("lCl", _I8), ("uCl", _I8),
        ("level_begin", _I8), ("level_size", _I8),
        ("Ceq", C.c_int32), ("deq", C.c_int32),
        ("Cin", C.c_int32), ("lA", C.c_int32), ("uA", C.c_int32),
        ("l", _I4), ("u", _I4),
        ("Ar", C.c_int32), ("br", C.c_int32), ("wr", C.c_int32),
        ("global_begin", C.c_int32), ("global_size", C.c_int32),
        ("stride", C.c_int32),
    ]


class Options(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("store_levels", C.c_int32),
                ("tol_feas", C.c_double), ("tol_rank", C.c_double)]


class EngineInfo(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("warps_per_cta", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("grid", C.c_int32),
                ("smem_per_cta", C.c_int64), ("blob_stride_bytes", C.c_int64),
                ("kernel_regs", C.c_int32), ("launches", C.c_int64)]


class GatherSrc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_stride", C.c_int64), ("rows", C.c_int32),
                ("cols", C.c_int32), ("src_ld", C.c_int32), ("row_index", C.c_void_p),
                ("dst_off", C.c_int32), ("scale", C.c_double)]


class PostureInputs(C.Structure):
    _fields_ = [("q", C.c_void_p), ("q_stride", C.c_int64), ("q_ref", C.c_void_p), ("q_ref_stride", C.c_int64),
                ("v_ref", C.c_void_p), ("v_ref_stride", C.c_int64),
                ("q_min", C.c_void_p), ("q_max", C.c_void_p), ("q_neutral", C.c_void_p), ("qdot_max", C.c_void_p),
                ("lambda_", C.c_double), ("bound_scaling", C.c_double), ("dT", C.c_double),
                ("postural_level", C.c_int32), ("joint_limit_set", C.c_int32),
                ("velocity_limit_set", C.c_int32), ("reserved", C.c_int32)]


def make_desc(n, levels, c_eq=0, c_in=0, n_bound_sets=1, eps_factor=2e2, m_reg=0,
              reg_weight_kind=W_IDENTITY):
    """levels: list of dicts {m, kind, wkind, has_c, c_loc}."""
    d = StackDesc()
    d.n, d.n_levels = n, len(levels)
    for k, lv in enumerate(levels):
        d.m[k] = lv.get("m", n)
        d.task_kind[k] = lv.get("kind", TASK_DENSE)
        d.weight_kind[k] = lv.get("wkind", W_IDENTITY)
        d.has_c[k] = lv.get("has_c", 0)
        d.c_loc[k] = lv.get("c_loc", 0)
    d.c_eq, d.c_in, d.n_bound_sets = c_eq, c_in, n_bound_sets
    d.m_reg, d.reg_weight_kind = m_reg, reg_weight_kind
    d.eps = eps_factor * OSB_QPOASES_BASE_EPS
    return d

"""Synthetic stacks for benchmarks and tests (include/opensot_b200_synth.h through the native
library): named shapes of SURVEY.md section 8 and seeded instance blobs."""
import ctypes as C

import numpy as np

from . import _abi
from .engine import lib, blob_layout


def _fns():
    L = lib()
    if not getattr(L, "_synth_ready", False):
        L.osb_synth_shape.argtypes = [C.c_char_p, C.c_double, C.POINTER(_abi.StackDesc)]
        L.osb_synth_fill.argtypes = [C.POINTER(_abi.StackDesc), C.c_uint64, C.c_uint64, C.c_int32, C.c_double,
                                     np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]
        L._synth_ready = True
    return L


def shape(name, eps_factor=2e2):
    d = _abi.StackDesc()
    if _fns().osb_synth_shape(name.encode(), eps_factor, C.byref(d)) != 0:
        raise ValueError("unknown shape " + name)
    return d


def fill(desc, seed, first, count, tight):
    """float64 blobs [count, stride] of instances first .. first+count-1."""
    stride = blob_layout(desc).stride
    blobs = np.zeros((count, stride), dtype=np.float64)
    rc = _fns().osb_synth_fill(C.byref(desc), seed, first, count, tight, blobs)
    if rc != 0:
        raise ValueError("bad descriptor")
    return blobs


def static_segments(desc, layout=None):
    """The fields of a synthetic blob that the generator never varies between instances or cycles (what a robot stack
    sets once: velocity limits, task weights, zero right-hand sides of contact rows, -inf sides of one-sided rows), as
    sorted (offset, count) pairs for BatchedStackSolver.wire_configure. A blob of the shape carries their values."""
    L = layout if layout is not None else blob_layout(desc)
    n = desc.n
    segs = []
    if desc.c_eq > 0:
        segs.append((L.deq, desc.c_eq))                      # Ceq x = 0
    if desc.c_in > 0:
        segs.append((L.lA, desc.c_in))                       # -inf <= Cin x
    if desc.n_bound_sets > 0:
        segs += [(L.l[0], n), (L.u[0], n)]                   # set 0: +-0.02 velocity limits
    for k in range(desc.n_levels):
        if desc.task_kind[k] == _abi.TASK_DENSE and desc.weight_kind[k] == _abi.W_DIAG:
            segs.append((L.w[k], desc.m[k]))                 # 0.1 / 1.0 task weights
        if desc.c_loc[k] > 0:
            segs.append((L.lCl[k], desc.c_loc[k]))
    segs.sort()
    merged = []
    for off, cnt in segs:
        if merged and merged[-1][0] + merged[-1][1] == off:
            merged[-1] = (merged[-1][0], merged[-1][1] + cnt)
        else:
            merged.append((off, cnt))
    return merged

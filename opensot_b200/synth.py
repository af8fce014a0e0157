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

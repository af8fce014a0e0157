"""TEST INFRASTRUCTURE: ctypes access to the CPU warp emulator of the solver kernels
(tests/emu/_bin/libosb_emu.so, built by build_emu.sh from the very .cuh sources nvcc compiles).
Never imported by anything under opensot_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

from opensot_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_bin", "libosb_emu.so")
GENERIC_F64, GENERIC_F32, RR_F64, RR_F32, MW_F64, MW_F32 = 0, 1, 2, 3, 4, 5
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("emu_driver.cpp", "warp_emu.h")]
    csrc = os.path.join(HERE, "..", "..", "opensot_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call([os.path.join(HERE, "build_emu.sh")])
    return SO


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_solve.argtypes = [C.POINTER(_abi.StackDesc), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.POINTER(_abi.Options), C.c_uint32, C.c_int, C.c_int,
                                C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def solve(desc, blobs, variant=RR_F64, order=0, levels=False, active_mask=0, max_iter=0):
    """blobs: fp64 [B, stride]. Returns x, status, x_levels (or None), iters, smem_bytes."""
    f32 = variant in (GENERIC_F32, RR_F32, MW_F32)
    dt = np.float32 if f32 else np.float64
    blobs = np.asarray(blobs, dtype=np.float64)
    B = blobs.shape[0]
    if f32:
        stride = (blobs.shape[1] + 3) & ~3
        b = np.zeros((B, stride), dtype=np.float32)
        b[:, :blobs.shape[1]] = blobs
    else:
        b = np.ascontiguousarray(blobs)
    x = np.zeros((B, desc.n), dtype=dt)
    st = np.zeros(B, dtype=np.int32)
    it = np.zeros(B, dtype=np.int32)
    xl = np.zeros((B, desc.n_levels, desc.n), dtype=dt) if levels else None
    opt = _abi.Options(max_iter, 1 if levels else 0, 0.0, 0.0)
    info = (C.c_int64 * 4)()
    rc = lib().emu_solve(C.byref(desc), b.ctypes.data, B, x.ctypes.data, st.ctypes.data,
                         xl.ctypes.data if levels else None, it.ctypes.data, C.byref(opt), active_mask,
                         variant, order, info)
    if rc == -1:
        return None
    assert rc == 0, rc
    return x, st, xl, it, int(info[0])

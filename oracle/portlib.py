"""ctypes access to the oracle (TEST INFRASTRUCTURE): the C restatement
(oracle/_port/libihqp_port.so) and the compiled reference (oracle/_ref/ref_driver).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs import this."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from opensot_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_port", "libihqp_port.so")
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def have_ref():
    return os.path.exists(REF_DRIVER)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PORT_SO):
            build_port()
        L = C.CDLL(PORT_SO)
        L.port_synth_shape.argtypes = [C.c_char_p, C.c_double, C.POINTER(_abi.StackDesc)]
        L.port_blob_layout.argtypes = [C.POINTER(_abi.StackDesc), C.POINTER(_abi.Layout)]
        L.port_synth_fill.argtypes = [C.POINTER(_abi.StackDesc), C.c_uint64, C.c_uint64,
                                      C.c_int32, C.c_double, _dp]
        L.port_solve_batch.argtypes = [C.POINTER(_abi.StackDesc), _dp, C.c_int32, C.c_void_p,
                                       _dp, C.c_void_p, _ip, _ip]
        L.port_cost_function.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_void_p,
                                         C.c_void_p, _dp, _dp]
        L.port_qp_solve.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, _dp,
                                    C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def shape(name, eps_factor=2e2):
    d = _abi.StackDesc()
    if lib().port_synth_shape(name.encode(), eps_factor, C.byref(d)) != 0:
        raise ValueError("unknown shape " + name)
    return d


def layout(desc):
    L = _abi.Layout()
    rc = lib().port_blob_layout(C.byref(desc), C.byref(L))
    if rc != 0:
        raise ValueError("bad descriptor (%d)" % rc)
    return L


def synth(desc, seed, first, count, tight):
    L = layout(desc)
    blobs = np.zeros((count, L.stride), dtype=np.float64)
    rc = lib().port_synth_fill(C.byref(desc), seed, first, count, tight, blobs)
    assert rc == 0
    return blobs


def port_solve(desc, blobs, active=None, levels=False):
    blobs = np.ascontiguousarray(blobs, dtype=np.float64)
    B = blobs.shape[0]
    x = np.zeros((B, desc.n))
    xl = np.zeros((B, desc.n_levels, desc.n)) if levels else None
    st = np.zeros(B, dtype=np.int32)
    it = np.zeros(B, dtype=np.int32)
    act = None
    if active is not None:
        act = np.ascontiguousarray(active, dtype=np.int32)
    rc = lib().port_solve_batch(C.byref(desc), blobs, B,
                                act.ctypes.data if act is not None else None, x,
                                xl.ctypes.data if levels else None, st, it)
    assert rc == 0
    return (x, xl, st, it) if levels else (x, st, it)


def _vp(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data


def port_qp(H, g, C_=None, lC=None, uC=None, l=None, u=None, eps=0.0):
    H = np.ascontiguousarray(H, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    n = g.size
    nc = 0 if C_ is None else np.asarray(C_).reshape(-1, n).shape[0]
    keep = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None
            for a in (C_, lC, uC, l, u)]
    x = np.zeros(n)
    it = C.c_int32(0)
    st = lib().port_qp_solve(n, H, g, nc, *[None if a is None else a.ctypes.data for a in keep],
                             eps, x, C.byref(it))
    return x, st, it.value


def _read_ref_out(path):
    hdr = np.fromfile(path, dtype=np.int32, count=4)
    n, L, count = int(hdr[0]), int(hdr[1]), int(hdr[2])
    d = np.fromfile(path, dtype=np.float64, offset=16).reshape(count, 1 + n * (1 + L))
    return d[:, 0] > 0.5, d[:, 1:1 + n].copy(), d[:, 1 + n:].reshape(count, L, n).copy()


def have_hcod():
    """third oracle: OpenSoT::solvers::HCOD over the vendored soth (make -C oracle hcod)"""
    return os.path.exists(REF_DRIVER + "_hcod")


def ref_solve(desc, blobs, backend="qpoases"):
    """Runs the UNMODIFIED reference on the given blobs: iHQP + back-end ("qpoases", "eiquadprog")
    or, with backend "hcod", OpenSoT::solvers::HCOD (level solutions are not reported then)."""
    if not have_ref():
        raise RuntimeError("oracle/_ref not built")
    if backend == "hcod" and not have_hcod():
        raise RuntimeError("oracle/_ref/ref_driver_hcod not built (make -C oracle hcod)")
    blobs = np.ascontiguousarray(blobs, dtype=np.float64)
    with tempfile.TemporaryDirectory() as td:
        dp, bp, op = (os.path.join(td, f) for f in ("desc.bin", "blobs.bin", "out.bin"))
        with open(dp, "wb") as f:
            f.write(bytes(desc))
        blobs.tofile(bp)
        subprocess.check_call([REF_DRIVER + ("_hcod" if backend == "hcod" else ""), "solveblob", dp, bp,
                               str(blobs.shape[0]), backend, op],
                              stdout=subprocess.DEVNULL)
        return _read_ref_out(op)


def ref_qp(k):
    """One BackEnd-level QP (dict with H,g,C,lC,uC,l,u,eps_factor) through the reference's
    qpOASES plugin. Returns x."""
    n = len(k["g"])
    nc = 0 if k.get("C") is None else np.asarray(k["C"]).reshape(-1, n).shape[0]
    parts = [np.array([k["eps_factor"]], dtype=np.float64),
             np.asarray(k["H"], dtype=np.float64).reshape(-1), np.asarray(k["g"], dtype=np.float64)]
    if nc:
        parts += [np.asarray(k["C"], dtype=np.float64).reshape(-1),
                  np.asarray(k["lC"], dtype=np.float64), np.asarray(k["uC"], dtype=np.float64)]
    parts += [np.asarray(k["l"], dtype=np.float64), np.asarray(k["u"], dtype=np.float64)]
    with tempfile.TemporaryDirectory() as td:
        ip = os.path.join(td, "qp.bin")
        with open(ip, "wb") as f:
            f.write(np.array([n, nc], dtype=np.int32).tobytes())
            f.write(np.concatenate(parts).tobytes())
        out = subprocess.check_output([REF_DRIVER, "qp", ip]).decode().strip().splitlines()[-1].split()
    if out[0] != "1":
        raise RuntimeError("reference back-end failed on KAT")
    return np.array([float(v) for v in out[1:]])


def ref_bench(shape_name, eps_factor, seed, count, tight, nproc, cycles, warmup=0):
    import json
    out = subprocess.check_output([REF_DRIVER, "bench", shape_name, repr(float(eps_factor)),
                                   str(seed), str(count), repr(float(tight)), str(nproc),
                                   str(cycles), str(warmup)])
    return json.loads(out.decode().strip().splitlines()[-1])


def _read_both(path):
    hdr = np.fromfile(path, dtype=np.int32, count=4)
    n, count = int(hdr[0]), int(hdr[2])
    d = np.fromfile(path, dtype=np.float64, offset=16).reshape(-1, 2 + 2 * n)
    return d[:, 0] > 0.5, d[:, 1] > 0.5, d[:, 2:2 + n].copy(), d[:, 2 + n:].copy()


def ref_solve_both(desc, blobs, nproc=None):
    """The UNMODIFIED reference (iHQP + qpOASES) on every blob, twice: cold (the constructor's
    initProblem cascade) and hot-started from that (a following solve(), the production path).
    Returns ok_hot, ok_cold, x_hot, x_cold. Work is split over `nproc` driver processes."""
    if not have_ref():
        raise RuntimeError("oracle/_ref not built")
    blobs = np.ascontiguousarray(blobs, dtype=np.float64)
    B = blobs.shape[0]
    if nproc is None:
        try:
            nproc = len(os.sched_getaffinity(0))
        except AttributeError:
            nproc = os.cpu_count() or 1
    nproc = max(1, min(nproc, B // 16 if B >= 16 else 1))
    cuts = [B * i // nproc for i in range(nproc + 1)]
    with tempfile.TemporaryDirectory() as td:
        dp = os.path.join(td, "desc.bin")
        with open(dp, "wb") as f:
            f.write(bytes(desc))
        procs = []
        for i in range(nproc):
            bp, op = os.path.join(td, "b%d.bin" % i), os.path.join(td, "o%d.bin" % i)
            blobs[cuts[i]:cuts[i + 1]].tofile(bp)
            procs.append((subprocess.Popen([REF_DRIVER, "solveboth", dp, bp, str(cuts[i + 1] - cuts[i]), op],
                                           stdout=subprocess.DEVNULL), op))
        parts = []
        for p, op in procs:
            if p.wait() != 0:
                raise RuntimeError("ref_driver solveboth failed")
            parts.append(_read_both(op))
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(4))


def ref_sequence(desc, blobs, cycles, seed=1):
    """`cycles` time-coherent control cycles (A_k += 0.01 N(0,1) per cycle) of
    persistent hot-started reference solvers, each next to a fresh (cold) solver on the same inputs.
    Returns blobs_seq [cycles, B, stride], ok_hot, ok_cold [cycles, B], x_hot, x_cold [cycles, B, n]."""
    if not have_ref():
        raise RuntimeError("oracle/_ref not built")
    blobs = np.ascontiguousarray(blobs, dtype=np.float64)
    B, stride = blobs.shape
    with tempfile.TemporaryDirectory() as td:
        dp, bp, op, sp = (os.path.join(td, f) for f in ("desc.bin", "blobs.bin", "out.bin", "seq.bin"))
        with open(dp, "wb") as f:
            f.write(bytes(desc))
        blobs.tofile(bp)
        subprocess.check_call([REF_DRIVER, "sequence", dp, bp, str(B), str(cycles), str(seed), op, sp],
                              stdout=subprocess.DEVNULL)
        okh, okc, xh, xc = _read_both(op)
        seq = np.fromfile(sp, dtype=np.float64).reshape(cycles, B, stride)
    n = desc.n
    return seq, okh.reshape(cycles, B), okc.reshape(cycles, B), xh.reshape(cycles, B, n), xc.reshape(cycles, B, n)


# ---- reference-defined parity set (used by the parity tests and by bench.py's parity block) ----
STABLE_TOL = 1e-8


def bounds_of(d, L, blobs):
    n = d.n
    lo = np.full((blobs.shape[0], n), -_abi.OSB_INFTY)
    hi = np.full((blobs.shape[0], n), _abi.OSB_INFTY)
    for s in range(d.n_bound_sets):
        lo = np.maximum(lo, blobs[:, L.l[s]:L.l[s] + n])
        hi = np.minimum(hi, blobs[:, L.u[s]:L.u[s] + n])
    return lo, hi


def max_violation(d, L, blobs, x):
    """largest violation of bounds / global rows by x (per instance)."""
    lo, hi = bounds_of(d, L, blobs)
    v = np.maximum(lo - x, x - hi).max(axis=1)
    n = d.n
    if d.c_in:
        Cin = blobs[:, L.Cin:L.Cin + d.c_in * n].reshape(-1, d.c_in, n)
        cx = np.einsum("bij,bj->bi", Cin, x)
        v = np.maximum(v, np.maximum(blobs[:, L.lA:L.lA + d.c_in] - cx, cx - blobs[:, L.uA:L.uA + d.c_in]).max(axis=1))
    if d.c_eq:
        Ceq = blobs[:, L.Ceq:L.Ceq + d.c_eq * n].reshape(-1, d.c_eq, n)
        cx = np.einsum("bij,bj->bi", Ceq, x)
        v = np.maximum(v, np.abs(cx - blobs[:, L.deq:L.deq + d.c_eq]).max(axis=1))
    return v


REF_FEAS_TOL = 1e-9


def live_stable_mask(d, L, blobs, ok_hot, ok_cold, x_hot, x_cold):
    """Reference-defined parity set for the LIVE reference: qpOASES hot-started and
    cold-started agree to 1e-8 AND the returned point satisfies the reference's own bounds and rows to
    1e-9. The second condition removes the cycles on which qpOASES stopped on its termination tolerance
    (1e9 EPS = 2.2e-7, external/qpOASES-ext/src/Options.cpp:206) outside the feasible set: measured
    here, every such answer (0-0.4 % of instances) is 1e-4..6e-4 from the optimum that eiQuadProg, HCOD
    and the restatement agree on to 1e-11, identically for hot and cold starts, while every answer that
    is feasible to 1e-9 is within 2e-10 of it. Returns (stable, hot_ne_cold, ref_infeasible)."""
    both = ok_hot & ok_cold
    differ = both & (np.abs(x_hot - x_cold).max(axis=-1) > STABLE_TOL)
    infeas = both & ~differ & (max_violation(d, L, blobs, x_hot) > REF_FEAS_TOL)
    return both & ~differ & ~infeas, differ, infeas


def ref_nhqp(desc, blobs, min_sv_ratio=0.05, ab_regularisation=True, ns_regularisation=True):
    """The reference's nHQP front-end (src/solvers/nHQP.cpp, unmodified) with qpOASES on every blob.
    Returns ok [B], x [B, n]."""
    if not have_ref():
        raise RuntimeError("oracle/_ref not built")
    blobs = np.ascontiguousarray(blobs, dtype=np.float64)
    B = blobs.shape[0]
    with tempfile.TemporaryDirectory() as td:
        dp, bp, op = (os.path.join(td, f) for f in ("desc.bin", "blobs.bin", "out.bin"))
        with open(dp, "wb") as f:
            f.write(bytes(desc))
        blobs.tofile(bp)
        subprocess.check_call([REF_DRIVER, "nhqp", dp, bp, str(B), repr(float(min_sv_ratio)),
                               "1" if ab_regularisation else "0", "1" if ns_regularisation else "0", op],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        d = np.fromfile(op, dtype=np.float64, offset=16).reshape(B, 1 + desc.n)
    return d[:, 0] > 0.5, d[:, 1:].copy()

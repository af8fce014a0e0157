"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/opensot_b200.h
declares, describes blobs consistently, and fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common
from opensot_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "opensot_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(osb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(native_lib):
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(native_lib, name), "libopensot_b200.so does not export " + name
    import opensot_b200.engine as E
    assert sorted(E.EXPORTS) == names, "python binding list out of sync with the header"


def test_struct_mirrors_match_the_c_side(native_lib):
    sizes = (C.c_int32 * 6)()
    native_lib.osb_struct_sizes(C.byref(sizes))
    mine = [C.sizeof(t) for t in (_abi.StackDesc, _abi.Layout, _abi.Options, _abi.EngineInfo, _abi.GatherSrc,
                                  _abi.PostureInputs)]
    assert list(sizes) == mine


def test_layout_of_named_shapes(native_lib, port):
    import opensot_b200 as ob
    for name, n, stride in [("C29", 29, 608), ("C35f", 35, 1160), ("C38c", 38, 1828)]:
        d = port.shape(name)
        L = ob.blob_layout(d)
        Lp = port.layout(d)
        assert bytes(L) == bytes(Lp), "product and oracle must describe the same blob"
        assert L.stride == stride and L.stride % 2 == 0
        # fields are disjoint and inside the blob
        spans = []
        for k in range(d.n_levels):
            if d.task_kind[k] == _abi.TASK_DENSE:
                spans.append((L.A[k], d.m[k] * n))
            spans.append((L.b[k], d.m[k]))
            if d.weight_kind[k] == _abi.W_DIAG:
                spans.append((L.w[k], d.m[k]))
        for s in range(d.n_bound_sets):
            spans += [(L.l[s], n), (L.u[s], n)]
        if d.c_eq:
            spans += [(L.Ceq, d.c_eq * n), (L.deq, d.c_eq)]
        if d.c_in:
            spans += [(L.Cin, d.c_in * n), (L.lA, d.c_in), (L.uA, d.c_in)]
        spans.sort()
        for (a, la), (b, lb) in zip(spans, spans[1:]):
            assert a >= 0 and a + la <= b
        assert spans[-1][0] + spans[-1][1] <= L.stride
        assert ob.smem_bytes_per_instance(d) > 8 * L.stride
    # algorithmic input size of SURVEY.md 8(d): 537 scalars for C29 (+ padding, + unit weights)
    d = port.shape("C29")
    assert 537 <= ob.blob_layout(d).stride <= 640


def test_invalid_descriptors_are_rejected(native_lib):
    import opensot_b200 as ob
    bad = [
        _abi.make_desc(0, [dict(m=1)]),
        _abi.make_desc(65, [dict(m=1)]),
        _abi.make_desc(5, []),
        _abi.make_desc(5, [dict(m=4, kind=_abi.TASK_IDENTITY)]),           # identity needs m == n
        _abi.make_desc(5, [dict(m=0, kind=_abi.TASK_HESSIAN), dict(m=5)]),  # H,g only on the last level
        _abi.make_desc(5, [dict(m=3)], n_bound_sets=5),
    ]
    for d in bad:
        with pytest.raises(ob.EngineError):
            ob.blob_layout(d)
    d = _abi.make_desc(5, [dict(m=3)])
    d.eps = -1.0
    with pytest.raises(ob.EngineError):
        ob.blob_layout(d)


def test_no_cpu_fallback(native_lib, port):
    """Without a CUDA device the engine refuses to exist; nothing routes through the oracle."""
    import opensot_b200 as ob
    if native_lib.osb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ob.EngineError, match="no CPU fallback"):
        ob.BatchedStackSolver(port.shape("C29"), 8)
    assert ob.measure_fp64_tflops(0, 16) < 0
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "opensot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ihqp_port" not in txt and "portlib" not in txt and "oracle/" not in txt.replace("oracle/ may", ""), f


def test_missing_library_raises(monkeypatch):
    import opensot_b200.engine as E
    monkeypatch.setattr(E, "_lib", None)
    monkeypatch.setattr(E, "_LIB_PATH", "/nonexistent/libopensot_b200.so")
    with pytest.raises(E.NativeLibraryMissing):
        E.lib()


def test_fp32_engine_rejects_explicit_rank_deficient_hessians(native_lib):
    """dense-weight / regularised levels with fewer task rows than variables form Hz explicitly: fp32
    cannot hold eps under |A|^2 u, the fp32 engine refuses them up front (OSB_E_PRECISION) instead of
    returning OSB_NOT_SPD per instance; the check precedes the device check (pure descriptor logic)"""
    import ctypes as C
    from opensot_b200 import synth, _abi
    for name, want in (("C29w", -5), ("C29r", -5)):
        d = synth.shape(name, _abi.F32_EPS_FACTOR)
        h = C.c_void_p()
        rc = native_lib.osb_engine_create_f32(C.byref(h), C.byref(d), 8, 0)
        assert rc == want, (name, rc, native_lib.osb_last_error())


def test_static_segments_of_the_synthetic_shapes(native_lib):
    """synth.static_segments (what bench.py hands to osb_wire_configure): sorted, disjoint, inside the blob, and the
    generator really never varies those fields -- over instances, seeds and the `tight` knob."""
    import opensot_b200 as ob
    from opensot_b200 import synth
    for name in ("C29", "C35f", "C38c", "C35h", "C64c", "C12", "C7c", "C16e"):
        d = synth.shape(name)
        L = ob.blob_layout(d)
        segs = synth.static_segments(d, L)
        assert segs == sorted(segs) and all(c > 0 for _, c in segs)
        assert all(a + ca <= b for (a, ca), (b, _) in zip(segs, segs[1:])) and segs[-1][0] + segs[-1][1] <= L.stride
        a = synth.fill(d, 1, 0, 16, 0.3)
        b = synth.fill(d, 99, 1000, 16, 0.05)
        for off, cnt in segs:
            assert np.all(a[:, off:off + cnt] == a[0, off:off + cnt]) and np.all(b[:, off:off + cnt] == a[0, off:off + cnt]), (name, off)
        dyn = np.ones(L.stride, dtype=bool)
        for off, cnt in segs:
            dyn[off:off + cnt] = False
        assert (a[:, dyn] != b[:, dyn]).any()

"""CPU suite, part 4: the CUDA kernels' LOGIC, run lane by lane on the CPU.

tests/emu compiles the very kernel sources nvcc compiles (opensot_b200/csrc/sot_kernel*.cuh)
with g++ and runs the 32 lanes of a warp as fibres (TEST INFRASTRUCTURE, see tests/emu/warp_emu.h;
nothing under opensot_b200/ loads it). What is checked here, without a GPU:
  * both solver kernels reproduce the oracle (same bar as the GPU parity tests, which remain the
    proof for the compiled SASS);
  * results are bit-identical for an ascending and a descending lane schedule with poisoned
    shared memory, i.e. no cross-lane shared-memory dependency is missing its __syncwarp() and
    nothing is read before it is written (this is how the y[ip] race of the generic kernel was
    found);
  * the host-side plan of the register-row kernel (shape -> kernel instance) is what DESIGN.md says."""
import ctypes as C
import os

import numpy as np
import pytest

import common
from opensot_b200 import _abi

emu = pytest.importorskip("emu.emulib")


@pytest.fixture(scope="module")
def emulib():
    emu.build()
    os.environ["EMU_POISON"] = "1"
    return emu


RR_CASES = [("C29", 0.1, 24), ("C29", 0.4, 12), ("C35f", 0.1, 10), ("C35h", 0.4, 10), ("C38c", 0.1, 8)]


@pytest.mark.parametrize("name,tight,count", RR_CASES)
def test_register_row_kernel_logic_matches_oracle(emulib, port, name, tight, count):
    d = port.shape(name)
    blobs = port.synth(d, 77, 0, count, tight)
    xp, xlp, stp, itp = port.port_solve(d, blobs, levels=True)
    runs = [emulib.solve(d, blobs, variant=emulib.RR_F64, order=o, levels=True) for o in (0, 1)]
    assert runs[0] is not None, "shape should be inside the register-row kernel's instances"
    x, st, xl, it, smem = runs[0]
    # same lane-order independent answer, bit for bit
    assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1])
    # same instances solved / flagged as the oracle, same answer
    assert np.array_equal(st != 0, stp != 0)
    ok = st == 0
    assert np.abs(x - xp)[ok].max() <= 1e-9
    assert np.abs(xl[:, -1] - xlp[:, -1])[ok].max() <= 1e-9
    L = port.layout(d)
    assert common.max_violation(d, L, blobs[ok], x[ok]).max() <= 1e-9


MW_CASES = [("C38c", 0.1, 8), ("C38c", 0.4, 6), ("C35f", 0.1, 8), ("C35h", 0.4, 8), ("C64c", 0.2, 4)]


@pytest.mark.parametrize("name,tight,count", MW_CASES)
def test_cta_per_stack_kernel_logic_matches_oracle(emulib, port, name, tight, count):
    """sot_kernel_mw.cuh (2-3 warps per stack, thread per row): the emulator runs the warps of the CTA as 64-96
    fibres, warp collectives rendez-vous per warp, __syncthreads over the CTA. Same bar as the register-row
    kernel; where both kernels take the shape, the answers are bit-identical (same arithmetic, other mapping);
    both thread orders over poisoned shared memory give the same bits (no missing barrier)."""
    d = port.shape(name)
    blobs = port.synth(d, 79, 0, count, tight)
    xp, xlp, stp, itp = port.port_solve(d, blobs, levels=True)
    runs = [emulib.solve(d, blobs, variant=emulib.MW_F64, order=o, levels=True) for o in (0, 1)]
    assert runs[0] is not None, "shape should be inside the CTA-per-stack kernel's instances"
    x, st, xl, it, smem = runs[0]
    assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1]) and np.array_equal(xl, runs[1][2])
    assert np.array_equal(st != 0, stp != 0)
    ok = st == 0
    assert np.abs(x - xp)[ok].max() <= 1e-9
    L = port.layout(d)
    assert common.max_violation(d, L, blobs[ok], x[ok]).max() <= 1e-9
    rr = emulib.solve(d, blobs, variant=emulib.RR_F64, levels=True)
    if rr is not None:
        assert np.array_equal(x, rr[0]) and np.array_equal(st, rr[1]) and np.array_equal(it, rr[3])
    # active-level masks (iHQP::setActiveStack) and the fp32 instance
    for active in ([1, 0, 1], [0, 1, 1]):
        mask = sum(1 << k for k, a in enumerate(active) if a)
        xa, sta = emulib.solve(d, blobs, variant=emulib.MW_F64, active_mask=mask)[:2]
        xpa, stpa, _ = port.port_solve(d, blobs, active=active)
        both = (sta == 0) & (stpa == 0)
        assert np.abs(xa - xpa)[both].max() <= 1e-9
    d32 = port.shape(name, _abi.F32_EPS_FACTOR)
    xf, stf = emulib.solve(d32, blobs, variant=emulib.MW_F32)[:2]
    xpf, stpf, _ = port.port_solve(d32, blobs)
    both = (stf == 0) & (stpf == 0)
    assert both.sum() >= count - 1 and np.abs(xf - xpf)[both].max() <= 1e-4


@pytest.mark.parametrize("name,tight,count", [("C7c", 0.3, 12), ("C12", 0.3, 12), ("C29", 0.1, 6)])
def test_generic_kernel_logic_matches_oracle(emulib, port, name, tight, count):
    d = port.shape(name)
    blobs = port.synth(d, 78, 0, count, tight)
    xp, xlp, stp, itp = port.port_solve(d, blobs, levels=True)
    runs = [emulib.solve(d, blobs, variant=emulib.GENERIC_F64, order=o, levels=True) for o in (0, 1)]
    x, st = runs[0][0], runs[0][1]
    assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1])
    assert (st == 0).all() and (stp == 0).all()
    assert np.abs(x - xp).max() <= 1e-9


@pytest.mark.parametrize("name,variant", [("C29w", "generic"), ("C29w", "rr"), ("C29r", "generic"), ("C20d", "generic"),
                                          ("C16e", "generic"), ("C29c", "rr")])
def test_every_level_path_logic(emulib, port, name, variant):
    """dense W, user regularisation task, weighted identity levels, equalities + inequalities +
    task-local rows (DENSE path: generic kernel), a dense weight whitened in place (register-row kernel: W = C C',
    the level becomes the identity-weight task (C'A, C'b)) and linear cost terms on LQ levels (register-row
    kernel): same instances flagged infeasible as the oracle, same answers, both lane orders."""
    d = port.shape(name)
    blobs = port.synth(d, 556, 0, 24, 0.2 if name.startswith("C29") else 0.3)
    xp, _, stp, _ = port.port_solve(d, blobs, levels=True)
    v = emulib.RR_F64 if variant == "rr" else emulib.GENERIC_F64
    runs = [emulib.solve(d, blobs, variant=v, order=o) for o in (0, 1)]
    assert runs[0] is not None
    x, st = runs[0][0], runs[0][1]
    assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1])
    assert np.array_equal(st != 0, stp != 0)
    ok = st == 0
    assert np.abs(x - xp)[ok].max() <= common.PARITY_TOL


def test_random_shapes_against_the_reference(emulib, port):
    """random stack shapes inside the register-row instances with a full-rank last level (well
    posed), arbitrated by the unmodified reference where oracle/_ref is present (else by the
    restatement): the kernel never fails where the reference solves, and agrees to the parity
    bound where both solve. tools/fuzz_shapes.py is the long version."""
    sys_path_tools = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools")
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_shapes", os.path.join(sys_path_tools, "fuzz_shapes.py"))
    fz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fz)
    rng = np.random.default_rng(11)
    checked = 0
    for it in range(12):
        d = fz.random_desc(rng, well_posed=True)
        blobs = port.synth(d, 900 + it, 0, 2, 0.2)
        r = emulib.solve(d, blobs, variant=emulib.RR_F64, order=it & 1)
        assert r is not None
        x, st = r[0], r[1]
        if port.have_ref():
            ok, xr, _ = port.ref_solve(d, blobs)
        else:
            xr, _, stp, _ = port.port_solve(d, blobs, levels=True); ok = stp == 0
        assert not ((st != 0) & ok).any(), (it, st, ok)
        both = (st == 0) & ok
        if both.any():
            assert np.abs(x - xr)[both].max() <= common.PARITY_TOL, it
            checked += int(both.sum())
    assert checked >= 12


def test_nan_anywhere_is_flagged(emulib, port):
    """a NaN in any field of an instance (level block, bound set, global row, its limits, an
    equality row) gives OSB_NAN_INPUT for that instance only (both kernels)"""
    for name in ("C38c", "C35f"):
        d = port.shape(name)
        L = port.layout(d)
        blobs = port.synth(d, 5, 0, 8, 0.1)
        spots = [L.A[1] + 3, L.b[2] + 1, L.l[1] + 2, L.u[0] + d.n - 1]
        if d.c_in:
            spots += [L.Cin + 5 * d.n + 3, L.uA + 2]
        if d.c_eq:
            spots += [L.Ceq + 2 * d.n + 1, L.deq + 3]
        for variant in (emulib.RR_F64, emulib.GENERIC_F64):
            for k, off in enumerate(spots):
                b = blobs.copy()
                b[k % 8, off] = np.nan
                x, st = emulib.solve(d, b, variant=variant)[:2]
                assert (st[k % 8] & 0xff) == _abi.NAN_INPUT, (name, variant, off)
                assert (np.delete(st, k % 8) == 0).all()


def test_register_row_kernel_logic_fp32(emulib, port):
    """stated fp32 tolerance of tests/test_gpu_parity.py: 1e-5 against the fp64 oracle"""
    d = port.shape("C29", _abi.F32_EPS_FACTOR)
    blobs = port.synth(d, 79, 0, 16, 0.1)
    xp, _, stp, _ = port.port_solve(d, blobs, levels=True)
    x, st, _, _, _ = emulib.solve(d, blobs, variant=emulib.RR_F32)
    assert (st == 0).all()
    assert np.abs(x - xp).max() <= 1e-5


def test_register_row_kernel_active_levels_and_golden(emulib, port):
    """setActiveStack semantics (iHQP.cpp:296-310) and a golden set of the unmodified reference"""
    d, z = common.load_golden("C29_t10")
    blobs = z["blobs"][:12]
    x, st, _, _, _ = emulib.solve(d, blobs, variant=emulib.RR_F64)
    xp, _, _, _ = port.port_solve(d, blobs, levels=True)
    stable = common.stable_mask(z, np.vstack([xp, np.zeros((len(z["blobs"]) - 12, d.n))]))[:12]
    assert (st == 0).all()
    assert np.abs(x - z["x_qpoases"][:12])[stable].max() <= common.PARITY_TOL
    d2 = port.shape("C29")
    b2 = port.synth(d2, 5, 0, 6, 0.2)
    for mask in (0b101, 0b110, 0b011):
        xa, sta, _, _, _ = emulib.solve(d2, b2, variant=emulib.RR_F64, active_mask=mask)
        xo, _, sto, _ = port.port_solve(d2, b2, active=[bool(mask >> k & 1) for k in range(3)], levels=True)
        # a rank-deficient last level is fixed by eps |x|^2 only: conditioned like 1/eps, hence the
        # parity bound of BASELINE.json rather than 1e-9
        assert (sta == 0).all() and np.abs(xa - xo).max() <= common.PARITY_TOL, mask


def test_plan_picks_the_documented_instances(native_lib):
    """shape -> kernel: n in [20,30] -> 30-column rows, n in (30,38] -> 38; one row per lane if
    every row count fits 32; narrow, wide, weighted-identity and user-regularised stacks -> generic; a dense weight on a
    dense task stays inside the instances (whitened in place)"""
    from oracle import portlib as P
    want = {"C29": (15, 1), "C35f": (19, 2), "C38c": (19, 2), "C35h": (19, 2), "C12": None, "C7c": None, "C64c": None}
    for name, inst in want.items():
        d = P.shape(name)
        info = (C.c_int64 * 4)()
        rc = emu.lib().emu_solve(C.byref(d), None, 0, None, None, None, None, None, 0, emu.RR_F64, 0, info)
        assert (rc == -1) == (inst is None), name
    dense = _abi.make_desc(29, [dict(m=6, wkind=_abi.W_DENSE), dict(kind=_abi.TASK_IDENTITY)])
    info = (C.c_int64 * 4)()
    assert emu.lib().emu_solve(C.byref(dense), None, 0, None, None, None, None, None, 0, emu.RR_F64, 0, info) != -1
    for other in (_abi.make_desc(29, [dict(m=6), dict(kind=_abi.TASK_IDENTITY, wkind=_abi.W_DIAG)]),      # weighted identity level
                  P.shape("C29r")):                                                                      # user regularisation task
        assert emu.lib().emu_solve(C.byref(other), None, 0, None, None, None, None, None, 0, emu.RR_F64, 0, None) == -1
    assert native_lib.osb_smem_bytes_per_instance(C.byref(P.shape("C29"))) <= 20 * 1024


@pytest.mark.parametrize("variant", ["f64", "f32"])
def test_projection_by_saved_reflectors(emulib, port, variant):
    """stacks with several dense levels and no global equalities: from the second dense level on
    the register-row kernel projects the task block with the saved reflectors of the earlier
    levels (rr_project_refl) while that is cheaper than the product with the basis, then falls
    back to the product -- both inside one stack here. Same answers as the oracle, both lane
    orders bit-identical, also with a level switched off (its reflectors do not exist)."""
    v = emulib.RR_F64 if variant == "f64" else emulib.RR_F32
    eps = 2e2 if variant == "f64" else _abi.F32_EPS_FACTOR
    tol = 1e-9 if variant == "f64" else 1e-4
    shapes = [_abi.make_desc(29, [dict(m=4), dict(m=5, wkind=_abi.W_DIAG), dict(m=6), dict(kind=_abi.TASK_IDENTITY)],
                             n_bound_sets=2, eps_factor=eps),
              _abi.make_desc(24, [dict(m=3), dict(m=12), dict(m=3), dict(kind=_abi.TASK_IDENTITY)], c_in=4, n_bound_sets=1,
                             eps_factor=eps)]     # full-rank last level: the answer is well defined to 1e-9
    for d in shapes:
        blobs = port.synth(d, 321, 0, 6, 0.2)
        for active in (None, [True, False] + [True] * (d.n_levels - 2)):
            xp, stp, _ = port.port_solve(d, blobs, active=active)
            mask = 0 if active is None else sum(1 << k for k, a in enumerate(active) if a)
            runs = [emulib.solve(d, blobs, variant=v, order=o, active_mask=mask) for o in (0, 1)]
            assert runs[0] is not None, "shape outside the register-row instances"
            x, st = runs[0][0], runs[0][1]
            assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1])
            assert np.array_equal(st != 0, stp != 0)
            ok = st == 0
            assert ok.any() and np.abs(x - xp)[ok].max() <= tol


def test_inactive_task_rows(emulib, port):
    """Task::setActive(false) leaves A = 0 and b unchanged (Task.h:382-386): zero task rows -- half of
    a level, a whole level -- are dependent rows to the sweep; same answers as the oracle (and as the
    live reference where present) from both kernels."""
    d = port.shape("C29"); L = port.layout(d); n = d.n
    blobs = port.synth(d, 77, 0, 8, 0.15)
    half = blobs.copy(); half[:, L.A[1] + 6 * n:L.A[1] + 12 * n] = 0.0
    whole = blobs.copy(); whole[:, L.A[1]:L.A[1] + 12 * n] = 0.0
    for bb in (half, whole):
        xp, stp, _ = port.port_solve(d, bb)
        assert (stp == 0).all()
        for v in (emulib.RR_F64, emulib.GENERIC_F64):
            r = emulib.solve(d, bb, variant=v)
            assert (r[1] == 0).all() and np.abs(r[0] - xp).max() <= 1e-9
        if port.have_ref():
            ok, xr, _ = port.ref_solve(d, bb)
            assert ok.all() and np.abs(xr - xp).max() <= common.STABLE_TOL


@pytest.mark.parametrize("name,tight,count,variant", [("C29c", 0.5, 24, "rr"), ("C29c", 0.8, 16, "rr"), ("C29", 0.8, 16, "rr"),
                                                      ("C35f", 0.8, 10, "mw"), ("C38c", 0.8, 8, "mw")])
def test_no_false_infeasible_at_degenerate_vertices(emulib, port, name, tight, count, variant):
    """Heavily saturated stacks and linear cost terms on rank-deficient levels end at fully degenerate vertices (more
    tight rows than free directions, 50-180 working-set changes). Round 1 reported OSB_INFEASIBLE there on instances the
    reference solves (a row violated by 1e-7..4e-6 of accumulated rounding with no step left). At a full vertex the
    kernels now recompute the iterate from the working set before believing a violation, and later levels' boxes contain
    the previous level's point: no instance fails that the live reference (iHQP+qpOASES) solves, answers stay feasible."""
    if not port.have_ref():
        pytest.skip("oracle/_ref not built")
    d = port.shape(name)
    blobs = port.synth(d, 4242, 0, count, tight)
    v = emulib.RR_F64 if variant == "rr" else emulib.MW_F64
    runs = [emulib.solve(d, blobs, variant=v, order=o) for o in (0, 1)]
    x, st = runs[0][0], runs[0][1]
    assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1])
    ok, xr, _ = port.ref_solve(d, blobs)
    assert not ((st != 0) & ok).any(), np.flatnonzero((st != 0) & ok)
    L = port.layout(d)
    assert common.max_violation(d, L, blobs[st == 0], x[st == 0]).max() <= 1e-9
    both = (st == 0) & ok
    assert np.median(np.abs(x - xr).max(axis=1)[both]) <= 1e-6


def test_dense_weight_whitened_on_the_cta_per_stack_kernel(emulib, port):
    """`W * task` with a matrix on a 38-DoF stack with 26 global rows (C38c with a dense level-1 weight): the CTA-per-stack
    kernel whitens the task in place (warp 0, W = C C') and stays on the LQ path; same answers as the oracle and as the
    generic kernel (which forms the reduced Hessian), both thread orders."""
    from opensot_b200 import _abi
    d = port.shape("C38c")
    d.weight_kind[1] = _abi.W_DENSE
    blobs = port.synth(d, 91, 0, 6, 0.1)
    xp, _, stp, _ = port.port_solve(d, blobs, levels=True)
    runs = [emulib.solve(d, blobs, variant=emulib.MW_F64, order=o) for o in (0, 1)]
    assert runs[0] is not None, "dense weight on a dense task belongs to the CTA-per-stack kernel's instances"
    x, st = runs[0][0], runs[0][1]
    assert np.array_equal(x, runs[1][0]) and np.array_equal(st, runs[1][1])
    assert np.array_equal(st != 0, stp != 0)
    ok = st == 0
    assert np.abs(x - xp)[ok].max() <= common.PARITY_TOL
    xg, stg = emulib.solve(d, blobs, variant=emulib.GENERIC_F64)[:2]
    assert np.array_equal(stg, st) and np.abs(x - xg)[ok].max() <= 1e-8


def test_dead_pair_skip_is_bit_identical(emulib, port, tmp_path):
    """The fp64 CTA-per-stack kernel enters its walks behind the pairs that lie left of the live basis (RR_KS_REFL,
    RR_KS_AS: exact zeros skipped under a warp-uniform branch). A second emulator build with the skip
    compiled out must give the same bits -- x, status, per-level solutions, iteration counts -- over the BASELINE shapes
    and random ones (20..38 and 39..64 columns, so that z0 falls on both sides of either threshold)."""
    import subprocess
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from fuzz_shapes import random_desc
    out = str(tmp_path / "noskip")
    env = dict(os.environ, EMU_OUT=out, EMU_NO_ABI="1", EMU_EXTRA="-DRR_KS_REFL=1000 -DRR_KS_AS=1000")
    subprocess.check_call([os.path.join(os.path.dirname(emu.__file__), "build_emu.sh")], env=env)
    plain = C.CDLL(os.path.join(out, "libosb_emu.so"))
    plain.emu_solve.argtypes = emulib.lib().emu_solve.argtypes

    def run(lib, d, blobs, variant, order):
        B = blobs.shape[0]
        f32 = variant == emulib.MW_F32
        dt = np.float32 if f32 else np.float64
        if f32:
            b = np.zeros((B, (blobs.shape[1] + 3) & ~3), dtype=np.float32); b[:, :blobs.shape[1]] = blobs
        else:
            b = np.ascontiguousarray(blobs)
        x = np.zeros((B, d.n), dtype=dt); st = np.zeros(B, dtype=np.int32); it = np.zeros(B, dtype=np.int32)
        xl = np.zeros((B, d.n_levels, d.n), dtype=dt)
        opt = _abi.Options(0, 1, 0.0, 0.0)
        info = (C.c_int64 * 4)()
        rc = lib.emu_solve(C.byref(d), b.ctypes.data, B, x.ctypes.data, st.ctypes.data, xl.ctypes.data, it.ctypes.data,
                           C.byref(opt), 0, variant, order, info)
        return None if rc == -1 else (x, st, xl, it)

    rng = np.random.default_rng(11)

    def wide(rng):
        n = int(rng.integers(39, 65)); nl = int(rng.integers(2, 4))
        levels = [dict(m=int(rng.integers(1, 16)), kind=_abi.TASK_DENSE, wkind=int(rng.integers(0, 2)), c_loc=0) for _ in range(nl - 1)]
        levels.append(dict(kind=_abi.TASK_IDENTITY, c_loc=0))
        return _abi.make_desc(n, levels, c_eq=int(rng.integers(0, 7)), c_in=int(rng.integers(0, 17)), n_bound_sets=1)

    descs = [port.shape(s) for s in ("C38c", "C35f", "C35h", "C64c")] + [random_desc(rng) if i % 3 else wide(rng) for i in range(16)]
    compared = skipped_something = 0
    for k, d in enumerate(descs):
        blobs = port.synth(d, 300 + k, 0, 3, float(rng.choice([0.1, 0.3, 0.6])))
        for variant in (emulib.MW_F64,):       # (the fp32 instances compile to the straight walks in both builds)
            a = run(emulib.lib(), d, blobs, variant, k & 1)
            b = run(plain, d, blobs, variant, k & 1)
            assert (a is None) == (b is None)
            if a is None:
                continue
            compared += 1
            for u, v in zip(a, b):
                assert np.array_equal(u, v, equal_nan=True), (k, variant)
        m = [d.m[j] for j in range(d.n_levels)]
        skipped_something += (d.c_eq + sum(m[:-1]) >= 8)      # some level starts with z0 >= 8: the reflect skip is taken
    assert compared >= 18 and skipped_something >= 8

"""CPU suite: the C++ host layer (opensot_b200/host: BatchediHQP, BatchedHCOD, the BackEnd plugin) against the
reference's own objects, with the CPU warp emulator of the kernels standing in for the GPU behind the C ABI
(tests/emu/emu_abi.cpp put in front of the real library with LD_LIBRARY_PATH; test infrastructure only).
The same binaries run over the real library on the GPU box in tests/test_gpu_host_cpp.py."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_bin", "test_dropin")
EMU_ABI = os.path.join(ROOT, "tests", "emu", "_bin", "abi")
LIBS = [EMU_ABI, os.path.join(ROOT, "opensot_b200", "lib"), os.path.join(ROOT, "oracle", "_ref")]


def check_surface(r, shape):
    """assertions shared with the GPU twin: every member of the iHQP / HCOD surface against the reference"""
    tol = 1e-6
    for k, v in r.items():
        if k.endswith("_err"):
            assert 0.0 <= v <= tol, (k, v)
        if k.endswith("_ok"):
            assert v == 1, k
    assert r["backend_rows_ok"] == 1 and r["last_level_is_x"] == 1 and r["range_checks"] == 1
    assert r["max_iter_hits"] > 0 and r["max_iter_solve_false"] == 1 and r["options_restored"] == 1
    assert r["eps_all_moved"] > 1e-6 and r["eps_level_moved"] > 1e-6      # the shifts mattered
    assert r["eps_rejects"] == 1 and r["mixed_box_active"] > 0
    assert r["identity_kind_before"] == 1 and r["identity_kind_after"] == 0     # OSB_TASK_IDENTITY -> OSB_TASK_DENSE
    # constraints whose row count differs between instances and changes per cycle (ConvexHull-like): solved through row slots,
    # the engine rebuilt exactly twice (when rows outgrew the slots), some of those rows binding at the optimum
    assert r["rows_change_rebuilds"] == 2 and r["rows_change_binding"] > 0
    assert r["rows_change_slots0"] == 5 and r["rows_change_slots"] == 13 and r["rows_change_min"] < r["rows_change_max"]
    assert r["log_entries"] >= 18      # Solver<>::log: the reference's variable names, sizes and values, level by level
    # a weight that gains off-diagonal terms / a linear cost term that appears after construction: rebuilt for the wider class
    assert r["weight_kind_before"] == 1 and r["weight_kind_after"] == 2 and r["has_c_before"] == 0 and r["has_c_after"] == 1
    assert r["weight_c_drift_moved"] > 1e-4        # the new W and c changed the reference's answer
    if shape == "C35f":
        assert r["eq_rows_before"] == 14 and r["eq_rows_after"] == 12


def run(args, libs):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(libs + [env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([BIN] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    res = {}
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            res.update(json.loads(line))
    return res


@pytest.fixture(scope="module")
def binaries():
    if not os.path.isdir("/root/reference/src") and not os.path.exists(BIN):
        pytest.skip("C++ drop-in binary not built (needs the OpenSoT/Eigen headers)")
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "opensot_b200", "host"), "USE_SHIMS=1"])
    from emu import emulib
    emulib.build()
    assert os.path.exists(os.path.join(EMU_ABI, "libopensot_b200.so"))


@pytest.mark.parametrize("shape,count", [("C29", 8), ("C38c", 6), ("C35f", 6), ("C35h", 6)])
def test_ihqp_and_hcod_surface_against_the_reference(binaries, shape, count):
    r = run(["surface", shape, count, 0.1, 31], LIBS)
    print(r)
    check_surface(r, shape)


def test_same_stacks_reference_batched_plugin(binaries):
    r = run(["C29", 12, 0.1, 31, 3], LIBS)
    print(r)
    assert r["ref_fail"] == 0 and r["ok_batched"] == 1 and r["bad_status"] == 0
    assert r["err_batched"] <= 1e-6 and r["plugin_fail"] == 0 and r["err_plugin"] <= 1e-6
    assert r.get("active_level_err", 0.0) <= 1e-6 and 0.0 <= r["err_hcod"] <= 1e-6


def check_nhqp(r):
    """BatchednHQP against the reference's own nHQP (src/solvers/nHQP.cpp through oracle/ref_nhqp.cpp): default
    settings and every switch flipped; each switch must move the answer (the second level carries a nearly
    dependent row, so the A/b damping is at work) and the joint velocities must agree to 1e-6"""
    for cfg in ("default", "no_ab", "no_ns", "sv20"):
        assert r[cfg + "_ok"] == 1 and r[cfg + "_ref_ok"] == r["count"], cfg
        assert 0.0 <= r[cfg + "_err"] <= 1e-6, (cfg, r[cfg + "_err"])
        if cfg != "default":
            assert r[cfg + "_moved"] > 1e-5, cfg
    assert r["single_ok"] == 1 and r["single_err"] <= 1e-6
    assert r["layers"][0] > r["layers"][1] > r["layers"][2] > 0


@pytest.mark.parametrize("shape,count", [("C29", 6), ("C38c", 4)])
def test_nhqp_front_end_against_the_reference_nhqp(binaries, shape, count):
    r = run(["nhqp", shape, count, 0.2, 31], LIBS)
    print(r)
    check_nhqp(r)

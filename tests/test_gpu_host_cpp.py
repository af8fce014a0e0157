"""GPU test of the C++ host surface: the same OpenSoT::AutoStack objects (built with the
reference's Math-of-Tasks operators) solved by (1) the reference iHQP+qpOASES, (2) our
BatchediHQP front-end, (3) the unmodified reference iHQP driving our BackEnd plugin through
BackEndFactory's dlopen (libOpenSotBackEndODYS.so). Binaries are built by
opensot_b200/host/Makefile where the OpenSoT/Eigen headers exist and travel with the repo."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_bin", "test_dropin")
LIBS = [os.path.join(ROOT, "opensot_b200", "lib"), os.path.join(ROOT, "oracle", "_ref")]


def run(shape, count, tight, seed, n_plugin):
    if not os.path.exists(BIN):
        if os.path.isdir("/root/reference/src"):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "opensot_b200", "host"), "USE_SHIMS=1"])
        else:
            pytest.skip("C++ drop-in binary not built (needs the OpenSoT/Eigen headers)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(LIBS + [env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([BIN, shape, str(count), repr(tight), str(seed), str(n_plugin)], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.strip().splitlines() if l.startswith("{")]
    res = {}
    for l in lines:
        res.update(l)
    print(res, out.stderr[-500:])
    return res


@pytest.mark.parametrize("shape,count,tight", [("C29", 96, 0.1), ("C35f", 48, 0.1), ("C38c", 32, 0.1), ("C35h", 32, 0.1)])
def test_same_cpp_stacks_reference_vs_batched_vs_plugin(shape, count, tight):
    r = run(shape, count, tight, 31, 6)
    assert r["ref_fail"] == 0
    assert r["ok_batched"] == 1 and r["bad_status"] == 0
    assert r["err_batched"] <= 1e-6          # joint-velocity infinity norm, fp64
    assert r["plugin_fail"] == 0 and r["err_plugin"] <= 1e-6
    assert r.get("active_level_err", 0.0) <= 1e-6
    assert 0.0 <= r["err_hcod"] <= 1e-6


@pytest.mark.parametrize("shape,count", [("C29", 64), ("C38c", 32), ("C35f", 32), ("C35h", 32)])
def test_ihqp_and_hcod_surface_against_the_reference(shape, count):
    """every constructor and member BatchediHQP / BatchedHCOD mirror (iHQP.h:59-193, HCOD.h:33-41), each against
    the reference object built with the same arguments; same assertions as the CPU twin over the emulator"""
    from test_host_cpp_emu import check_surface
    if not os.path.exists(BIN):
        pytest.skip("C++ drop-in binary not built (needs the OpenSoT/Eigen headers)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(LIBS + [env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([BIN, "surface", shape, str(count), "0.1", "31"], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    r = {}
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            r.update(json.loads(line))
    print(r)
    check_surface(r, shape)


def test_host_path_throughput_is_reported():
    """tests/cpp/bench_dropin: updateAndSolve for B = 4096 against the reference loop on the same AutoStack objects,
    and the batch-of-one latencies; the numbers land in DESIGN.md, here only sanity and parity are asserted"""
    bench = os.path.join(ROOT, "tests", "cpp", "_bin", "bench_dropin")
    if not os.path.exists(bench):
        pytest.skip("bench_dropin not built")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(LIBS + [env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([bench, "C29", "1024", "0.1", "31", "5"], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    print(r)
    assert r["ok"] == 1 and r["ref_fail"] == 0 and r["err_vs_reference"] <= 1e-6 and r["plugin_ok"] == 1
    assert r["batched_stacks_per_s"] > r["reference_stacks_per_s_1thread"]


@pytest.mark.parametrize("shape,count", [("C29", 64), ("C38c", 32), ("C12", 64)])
def test_nhqp_front_end_against_the_reference_nhqp(shape, count):
    """the nHQP recursion (per-level SVD with A/b damping and selective null-space regularisation on
    host threads, every level's B QPs in one C-ABI call) against the unmodified src/solvers/nHQP.cpp"""
    from test_host_cpp_emu import check_nhqp
    if not os.path.exists(BIN):
        pytest.skip("C++ drop-in binary not built (needs the OpenSoT/Eigen headers)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join(LIBS + [env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([BIN, "nhqp", shape, str(count), "0.2", "31"], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    r = {}
    for line in out.stdout.splitlines():
        if line.startswith("{"):
            r.update(json.loads(line))
    print(r)
    check_nhqp(r)

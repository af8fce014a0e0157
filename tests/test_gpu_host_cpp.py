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
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "opensot_b200", "host")])
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

"""The solve fused with the all-gather over peer memory (osb_gather_*): two
ranks, each solving half of a batch, must both end up with the whole batch's solutions, equal to
the unfused path bit for bit and to the oracle within the parity tolerance. On a box with one
GPU both ranks share it (CUDA IPC between two processes of one device)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape,batch,prec", [("C29", 1024, "f64"), ("C12", 515, "f64"), ("C29", 512, "f32")])
def test_fused_solve_allgather_two_ranks(shape, batch, prec):
    tol = 1e-6 if prec == "f64" else 1e-4      # BASELINE.json's parity bound; fp32 as in test_gpu_parity.py
    env = dict(os.environ, OSB_SHAPE=shape, OSB_BATCH=str(batch), OSB_PRECISION=prec, OSB_STEPS="0", OSB_TOL=str(tol))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "fused_gather_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-4000:]
    out = json.loads(lines[-1])
    assert out["ok_all_ranks"] and out["barrier_failures"] == 0 and out["equal_to_unfused"], out
    assert out["status_agree"] and out["max_err_vs_oracle"] <= tol, out

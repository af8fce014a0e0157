"""CPU suite: the bench.py contract. The reference arm (`--impl reference`: the unmodified reference,
or the restatement when oracle/_ref is absent, on the host cores) runs here and prints the agreed
JSON line; the committed bench lines of this round carry every key the contract names."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_runs_on_the_host(port):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and BASE_KEYS <= set(line)
    assert line["value"] > 0 and line["warmup"] >= 3 and line["dtype"] == "f64"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    # same workload as our arm at the same flags (the driver compares the two `config` objects)
    import bench
    args = type("A", (), dict(shape="C29", batch=4096, strong=False))()
    assert line["config"] == bench.config_of(args, 1)


def test_work_per_stack_reproduces_survey_constants():
    """algorithmic flops and bytes per stack of the three named shapes"""
    import bench
    assert bench.work_per_stack("C29") == (319577.0, 4528.0)
    assert bench.work_per_stack("C35f") == (370193.0, 8992.0)
    assert bench.work_per_stack("C38c") == (712118.0, 14216.0)


@pytest.mark.parametrize("name", ["r01_bench_1gpu.json", "r01_bench_2gpu.json", "r01_bench_4gpu.json", "r01_bench_8gpu.json"])
def test_committed_bench_lines_carry_the_contract(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    for key in ("roofline", "clocks", "gpu_launches"):
        assert key in line, key
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert line["gpu_launches"] == line["steps"] and line["unsolved_instances"] == 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < line["value"]
    if line["n_gpus"] == 1:
        assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] > 0
    else:
        assert line["fused_allgather"]["barrier_failures"] == 0 and line["fused_allgather"]["matches_unfused"]

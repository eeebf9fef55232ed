"""bench.py's reference arm runs without a GPU (it times the CPU oracle), so its JSON line -- the
same contract as the GPU arm's, with `impl: reference` -- can be checked here: required keys,
the metric of BASELINE.json, and the same `config` object the GPU arm prints for that workload
(the driver compares the two)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,keyframes", [("os1-64", 8), ("vlp-16", 12)])
def test_reference_arm_line(workload, keyframes):
    d = _line("--workload", workload, "--keyframes", str(keyframes), "--steps", "1", "--warmup", "1")
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "keyframes/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["value"] > 0
    assert d["config"]["workload"] == workload and d["config"]["keyframes_per_step_per_gpu"] == keyframes
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if workload == "os1-64":
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            assert d["metric"] == json.load(f)["metric"]


def test_both_arms_describe_the_workload_identically():
    """The `config` object comes from one function for both arms (bench.workload_config)."""
    sys.path.insert(0, ROOT)
    import bench
    import inspect
    src = inspect.getsource(bench)
    assert src.count("workload_config(args, p, ") >= 2  # the GPU arm and cpu_reference_arm
    assert set(bench.WORKLOADS) == {"os1-64", "vlp-16", "os1-64-dense", "os1-128", "assoc-100k"}
    assert bench.WORKLOADS["os1-64"][2] == bench.METRIC_OS1_64

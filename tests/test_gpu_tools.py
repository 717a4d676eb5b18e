"""The benchmark / profile drivers of SURVEY.md 8(f)-4 (Python twins of the reference's benchmark/*.jl scripts) run
end to end on the GPU box: they are user-facing entry points, so a broken driver is a broken deliverable even when
every kernel is right."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    r = subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
    return r.stdout


def test_benchmark_sweep_runs_the_six_reference_configurations():
    """benchmark/deposit_benchmark.jl + full_pipeline_benchmark.jl: six configurations, GPU side"""
    out = _run("tools/benchmark_sweep.py", "--no-cpu", "--reps", "2", "--deposit")
    assert out.count("^3") >= 12                      # six rows per table
    with open(os.path.join(ROOT, "gpurun_out", "benchmark_sweep.json")) as f:
        rep = json.load(f)
    assert len(rep["rows"]) == 6 and len(rep["deposit_only"]) == 6
    for row in rep["rows"]:
        assert all(row["gpu_ms"][k] > 0 for k in ("deposit", "solve", "interpolate", "pipeline"))
    for row in rep["deposit_only"]:                   # the reference prints the charge-conservation error in percent
        assert row["charge_error_percent"] < 1e-8


def test_full_pipeline_profile_gpu_flag():
    """benchmark/full_pipeline_profile.jl --gpu: one cold and one warm pipeline with per-pass times"""
    out = _run("tools/full_pipeline_profile.py", "--gpu", "--workload", "basic")
    rep = json.loads(out[out.index("{"):])
    assert rep["backend"] == "b200" and rep["particles"] == 100000
    assert rep["cold_geometry"]["green_build_ms"] > 0 and rep["warm"]["green_build_ms"] == 0
    assert set(rep["warm"]["passes_ms"]) == {"F1", "F2", "Z", "B2", "B3"} and rep["warm"]["launches"] >= 7


def test_stage_only_bench_line():
    out = _run("bench.py", "--stages-only", "--workload", "basic")
    line = json.loads(out.strip().splitlines()[-1])
    assert line["workload"] == "basic" and all(v > 0 for v in line["stages_ms"].values())

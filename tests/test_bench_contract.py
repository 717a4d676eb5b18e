"""bench.py's reference arm as the driver launches it (no GPU needed): one JSON line from rank 0 with the
contract's keys, the other ranks exit 0 without work; and the byte model behind `roofline.achieved`
against SURVEY.md 8(d)'s figures for config 5."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(cmd):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def _check(line):
    assert line["impl"] == "reference" and line["unit"] == "particles/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("basic")


def test_reference_arm_single_process():
    _check(_reference_line([sys.executable, "bench.py", "--impl", "reference", "--workload", "basic", "--steps", "2",
                            "--warmup", "1"]))


def test_reference_arm_under_torchrun_two_ranks():
    port = 29600 + (os.getpid() % 300)
    line = _reference_line([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", str(port), "bench.py", "--impl", "reference",
                            "--gpus", "2", "--workload", "basic", "--steps", "1", "--warmup", "0"])
    _check(line)
    assert line["n_gpus"] == 2


def test_algorithmic_bytes_match_the_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    ab = bench.algorithmic_bytes(100_000_000, (256, 256, 256), 8, False)
    # SURVEY.md 8(d), config 5 Float64, in MB
    want = {"deposit": 3334.2, "interpolate": 5202.7, "F1": 403.7, "F2": 808.5, "Z": 2563.3, "B2": 2425.4, "B3": 1211.1}
    for k, mb in want.items():
        assert abs(ab[k] / 1e6 - mb) < 0.06, (k, ab[k] / 1e6, mb)
    ab3 = bench.algorithmic_bytes(10_000_000, (128, 128, 128), 8, False)
    assert abs(sum(ab3[k] for k in ("F1", "F2", "Z", "B2", "B3")) / 1e6 - 930.2) < 0.1
    cath = bench.algorithmic_bytes(10_000_000, (128, 128, 256), 8, True)
    assert abs(cath["Z"] / 1e6 - 1052.7) < 0.1

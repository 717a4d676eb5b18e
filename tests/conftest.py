import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def scb():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    from oracle import spacecharge_oracle
    return spacecharge_oracle


_REPORT = []


@pytest.fixture
def record(request):
    """record(name, value, tol): collect measured parity errors; dumped to gpurun_out/ at the end."""
    def _rec(name, value, tol=None):
        _REPORT.append({"test": request.node.name, "what": name, "err": float(value), "tol": tol})
    return _rec


def pytest_sessionfinish(session, exitstatus):
    if not _REPORT:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(_REPORT, f, indent=1)

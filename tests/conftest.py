import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_golden.json")) as f:
        return json.load(f)


# Every parity comparison that needed the "band the reference spans over rank counts" instead of the plain 2 % / 1e-7 criteria is
# recorded here by tests/test_gpu_parity.py::check_parity and listed at the end of the run (and in gpurun_out/parity_band_report.json).
BAND_REPORT = []


def pytest_terminal_summary(terminalreporter):
    if not BAND_REPORT:
        return
    terminalreporter.section("parity comparisons that needed the oracle's rank-count band")
    for e in BAND_REPORT:
        terminalreporter.write_line(str(e))
    try:
        import json
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_band_report.json"), "w") as f:
            json.dump(BAND_REPORT, f, indent=1)
    except Exception:
        pass

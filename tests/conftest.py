import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle and the product library exist (incremental make)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    if not os.path.exists(os.path.join(ROOT, "nraps_b200", "lib", "libnraps_b200.so")):
        subprocess.run(["make", "-j", "8", "-C", os.path.join(ROOT, "nraps_b200", "csrc")], check=True, capture_output=True)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "pyref: needs the scratch build of the Python reference (oracle/build_pyref.py)")


@pytest.fixture(scope="session")
def oracle_libs():
    """Build the CPU oracle (port) if needed; the reference-derived library is prebuilt."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return os.path.join(ROOT, "oracle")

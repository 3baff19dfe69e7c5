import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The CPU-side libraries must exist (built by `make` / __graft_entry__.build())."""
    import subprocess
    need = [os.path.join(ROOT, "lighthugger_b200", "lib", "liblh_host.so"), os.path.join(ROOT, "oracle", "liblh_oracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT, "host", "oracle"])

"""pytest configuration: markers, import path and lazily built native libraries.

`-m "not gpu"` runs on a CPU-only box: the oracle against the golden vectors,
host logic, and the C-ABI export check.  `-m gpu` are the parity tests proper
and call the CUDA kernels through the C ABI.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _make(directory):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, directory)], check=True)


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Build whatever is missing (sources only are in git; .so files are ignored)."""
    if not (os.path.exists(os.path.join(ROOT, "oracle", "libgjk_oracle_f64.so"))
            and os.path.exists(os.path.join(ROOT, "oracle", "libgjk_oracle_f32.so"))):
        _make("oracle")
    if not os.path.exists(os.path.join(ROOT, "opengjk_b200", "libopengjk_b200.so")):
        _make(os.path.join("opengjk_b200", "csrc"))
    yield


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False

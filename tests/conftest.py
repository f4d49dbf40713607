import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "av1-go_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def built():
    """Make sure libav1r.so and the oracle library exist (nvcc cross-compiles without a GPU)."""
    import subprocess
    lib = os.path.join(ROOT, "av1-go_b200", "lib", "libav1r.so")
    orc = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        subprocess.check_call(["make", "-j8"], cwd=ROOT)
    return lib, orc

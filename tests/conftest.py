"""pytest configuration.

`-m "not gpu"` (run by the driver in the GPU-less authoring container) covers the oracle against the
reference build and the golden vectors, the host logic, the CPU emulation of the device code and the
C-ABI export check.  `-m gpu` holds the parity tests proper: they call through the C ABI of
brutefir_b200/libbfcuda.so and compare with the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_libs():
    """Build (if needed) and load the CPU checkers.  `ref` is None when oracle/_ref was never built."""
    from oracle import pyoracle
    if not pyoracle.available("oracle") or (os.path.isdir("/root/reference") and not pyoracle.available("ref")):
        pyoracle.build()
    return {"oracle": pyoracle.lib("oracle"), "ref": pyoracle.lib("ref") if pyoracle.available("ref") else None}


@pytest.fixture(scope="session")
def gpu_lib():
    from brutefir_b200 import _abi
    lib = _abi.load_library()      # raises loudly if the extension is missing: no fallback
    if lib.bfcuda_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible")
    return lib

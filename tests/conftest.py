import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libgspb200_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY CPU-emulated build of the same kernels (tests/emu); never used by the product."""
    subprocess.run(["make", "-C", EMU_DIR, "-j8"], check=True, stdout=subprocess.DEVNULL)
    import gsp_b200 as gsp

    lib = gsp.Library(EMU_LIB)
    assert "EMULATION" in lib.version()
    yield lib
    lib.close()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on cuda:0.  Fails (not skips) if the extension is missing."""
    import gsp_b200 as gsp

    lib = gsp.Library()
    assert "sm_100a" in lib.version()
    yield lib
    lib.close()

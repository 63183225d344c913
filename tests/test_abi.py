"""The C-ABI shared library loads and exports every symbol include/gsp_b200.h declares (no compute
without a GPU), and the product path fails loudly - no fallback - when no CUDA device is usable."""
import ctypes
import os
import re
import subprocess

import pytest

import gsp_b200 as gsp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gsp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsp_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product_so():
    subprocess.run(["make", "-C", os.path.join(ROOT, "geostatsprocesses.jl_b200", "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    return ctypes.CDLL(gsp.DEFAULT_LIB)


def test_header_symbols_exported(product_so):
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(product_so, s), f"{s} declared in gsp_b200.h but not exported"


def test_binding_covers_header():
    assert sorted(gsp.SIGNATURES) == declared_symbols()


def test_product_is_sm100a_only(product_so):
    product_so.gsp_version.restype = ctypes.c_char_p
    assert b"sm_100a" in product_so.gsp_version()
    out = subprocess.run(["cuobjdump", "-lelf", gsp.DEFAULT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device failure cannot be exercised")
    with pytest.raises(gsp.GspError):
        gsp.Library()  # gsp_ctx_create -> GSP_E_CUDA, nothing else is substituted


def test_missing_library_raises(tmp_path):
    with pytest.raises(FileNotFoundError):
        gsp.Library(str(tmp_path / "libgspb200.so"))


def test_argument_errors_emulated(emu_lib):
    """Error behaviour of the boundary (-k = argument k), exercised on the emulated build."""
    import numpy as np
    from helpers import iso
    import gsp_oracle as O

    st = iso(O.SPHERICAL, 1.0, 5.0, 2)
    dom = (gsp._lib.make_grid_domain((8, 8), (0, 0), (1, 1)), None)
    with pytest.raises(ValueError):  # descending dinds (must be findall(mask) order)
        gsp.LUPlan(emu_lib, st, dom, np.array([5, 3]), np.array([0.0, 1.0]), 0.0)
    with pytest.raises(ValueError):  # out of range
        gsp.LUPlan(emu_lib, st, dom, np.array([65]), np.array([0.0]), 0.0)
    with pytest.raises(ValueError):  # unknown model kind
        gsp.LUPlan(emu_lib, [(17, 1.0, np.eye(3))], dom, None, None, 0.0)
    with pytest.raises(gsp.GspError):  # a prime extent whose line does not fit shared memory: unsupported, reported not faked
        gsp.FFTPlan(emu_lib, st, (20011, 2), (0, 0), (1, 1))
    plan = gsp.FFTPlan(emu_lib, st, (8, 8), (0, 0), (1, 1))
    with pytest.raises(ValueError):
        plan.sample(1, np.zeros((1, 64)) + 0.5, sill=-1.0)
    with pytest.raises(ValueError):
        plan.sample(1, np.zeros((1, 64)) + 0.5, inds1=np.array([0, 3]))
    lp = gsp.LUPlan(emu_lib, st, dom, None, None, 0.0)
    with pytest.raises(ValueError):  # W1 without W: both noises are injected or both come from the device RNG
        lp.sample(2, None, rho=0.5, W1=np.zeros((lp.Ns, 2)))
    lp.close()
    # non-positive-definite matrix -> PosDefException(info) like cholesky (lusim.jl:92)
    A = np.eye(5)
    A[3, 3] = -1.0
    with pytest.raises(gsp.PosDefException) as ei:
        emu_lib.potrf(A)
    assert ei.value.info == 4

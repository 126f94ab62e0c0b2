"""pytest configuration.

Two tiers:
  -m "not gpu"  CPU only: the oracle against the reference's golden vectors, the product's host logic (compiled
                for the CPU from the same headers the kernels use), C-ABI symbol checks, world_size-2 gloo test.
  -m gpu        the parity tests proper: every one calls the CUDA path through the C ABI (libppcr_cuda.so) and
                compares with the oracle on the same seeded inputs.  They FAIL (not skip) without a B200.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200; runs the CUDA path through the C ABI")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def emu():
    """CPU build of the product's host/device-shared headers (tests/emu)."""
    import ctypes as C
    here = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(here, "libppcr_emu.so")
    src = os.path.join(here, "emu_host_logic.cpp")
    hdrs = [os.path.join(ROOT, "probabilistic_point_clouds_registration_b200", "csrc", h) for h in ("ppcr_lm.h", "ppcr_eval.h", "ppcr_tree.h")]
    if (not os.path.exists(so)) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in [src, *hdrs]):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


@pytest.fixture(scope="session")
def capi():
    from probabilistic_point_clouds_registration_b200 import build, capi as K
    build.build_cuda()
    K.lib()
    return K

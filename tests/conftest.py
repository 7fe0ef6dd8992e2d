import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def hostemu():
    """Test-only host emulation of the kernel bodies (tests/hostemu/hostemu.cpp), built on demand with g++."""
    so = os.path.join(ROOT, "tests", "hostemu", "libvmsm_hostemu.so")
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    csrc = os.path.join(ROOT, "verifiable_mpc_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.hostemu_msm.restype = ctypes.c_uint32
    lib.hostemu_choose_window.restype = ctypes.c_uint32
    return lib


@pytest.fixture(scope="session")
def ctx():
    """A live engine context on cuda:0 -- fails loudly (no fallback) if the extension or the GPU is missing."""
    from verifiable_mpc_b200 import Context

    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def known_points():
    """300 oracle-generated points with known discrete logs: g_i = scalar(0x5EEE, i) * B."""
    from oracle import ed25519 as E
    from oracle import prng

    dl = [prng.scalar(0x5EEE, i) for i in range(300)]
    return dl, [E.scalar_mul(E.B, r) for r in dl]

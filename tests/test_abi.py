"""The C-ABI library loads without a GPU, exports every symbol include/vmsm.h declares, the ctypes table mirrors
the header, and the product fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "vmsm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vmsm_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_ctypes_table():
    from verifiable_mpc_b200 import _lib

    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_symbol():
    from verifiable_mpc_b200 import _lib

    lib = _lib.load()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.vmsm_version() >= 100


def test_no_cpu_fallback_and_no_oracle_import():
    import subprocess
    import sys

    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import verifiable_mpc_b200 as v\n"
        "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'product imported oracle'\n"
        "import ctypes\n"
        "n = ctypes.c_int32()\n"
        "rc = v._lib.load().vmsm_device_count(ctypes.byref(n))\n"
        "if rc == 0 and n.value > 0:\n"
        "    print('HAS_GPU')\n"
        "else:\n"
        "    try:\n"
        "        v.Context(0)\n"
        "    except v.VmsmError as e:\n"
        "        print('LOUD', e.code)\n"
    ) % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout
    assert "HAS_GPU" in out or "LOUD -2" in out, out


def test_product_sources_never_reference_oracle():
    pkg = os.path.join(ROOT, "verifiable_mpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f

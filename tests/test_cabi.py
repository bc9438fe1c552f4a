"""The C-ABI library loads and exports everything include/b2sv.h declares (CPU only, no compute)."""

import ctypes
import os
import re

import numpy as np
import pytest

from unitaria_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2sv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2sv_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/b2sv.h but not exported"
        assert name in cabi.SYMBOLS, f"{name} has no ctypes prototype in unitaria_b200/cabi.py"
    assert set(cabi.SYMBOLS) == set(names)


def test_abi_version_and_struct_layouts():
    lib = cabi.load()
    assert lib.b2sv_abi_version() == 2
    assert cabi.GATE_DTYPE.itemsize == 80 and cabi.SUBINSTR_DTYPE.itemsize == 32
    assert ctypes.sizeof(cabi.PlanOpts) == 32
    assert ctypes.sizeof(cabi.Stats) == 8 * 3 + 8 * 12 * 3 + 8


def test_no_gpu_fails_loudly_not_silently():
    """Without a CUDA device the engine refuses to compute (there is no CPU fallback)."""
    if cabi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from unitaria_b200 import StateVector

    with pytest.raises(cabi.B2svError) as e:
        StateVector(3)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_argument_validation_needs_no_device():
    lib = cabi.load()
    need = ctypes.c_size_t()
    rc = lib.b2sv_plan_describe(0, 0, None, 0, None, None, 0, ctypes.byref(need))
    assert rc == -1 and b"n_qubits" in lib.b2sv_last_error()
    gates = np.zeros(1, dtype=cabi.GATE_DTYPE)
    assert lib.b2sv_plan_describe(3, 0, gates.ctypes.data, 1, None, None, 0, ctypes.byref(need)) == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "unitaria_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "sv_oracle" not in text and "np_oracle" not in text, f

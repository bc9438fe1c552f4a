"""ctypes wrapper for oracle/sv_oracle.c.  TEST INFRASTRUCTURE ONLY (see sv_oracle.c header)."""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsv_oracle.so")

KIND = {"I": 0, "X": 1, "Y": 2, "Z": 3, "H": 4, "RX": 5, "RY": 6, "RZ": 7, "PHASE": 8, "GLOBALPHASE": 9, "SWAP": 10}

GATE_DTYPE = np.dtype(
    [("kind", np.int32), ("t0", np.int32), ("t1", np.int32), ("_pad", np.int32), ("ctrl_mask", np.uint64), ("param", np.float64)]
)
assert GATE_DTYPE.itemsize == 32

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sv_oracle.c")
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_apply.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64]
        _lib.orc_apply.restype = None
        _lib.orc_set_basis.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64]
        _lib.orc_set_basis.restype = None
        _lib.orc_project_norm2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.orc_project_norm2.restype = ctypes.c_double
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def encode_gates(gates) -> np.ndarray:
    """tequila-like gate records -> the oracle's own wire format (struct orc_gate)."""
    out = np.zeros(len(gates), dtype=GATE_DTYPE)
    for j, g in enumerate(gates):
        name = g.name.upper()
        out["kind"][j] = KIND[name]
        tgt = tuple(g.target or ())
        out["t0"][j] = tgt[0] if len(tgt) > 0 else -1
        out["t1"][j] = tgt[1] if len(tgt) > 1 else -1
        m = 0
        for c in g.control or ():
            m |= 1 << int(c)
        out["ctrl_mask"][j] = m
        if name in ("X", "Y", "Z", "H", "SWAP", "I"):
            p = getattr(g, "power", 1)
            out["param"][j] = 1.0 if p is None else float(p)
            if name == "SWAP":
                assert out["param"][j] == 1.0
        else:
            out["param"][j] = float(g.parameter)
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> None:
    """Use ``n`` OpenMP threads (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().orc_set_threads(int(n))


def apply_encoded(psi: np.ndarray, n: int, enc: np.ndarray) -> None:
    assert psi.dtype == np.complex128 and psi.flags.c_contiguous and psi.size == 2**n
    enc = np.ascontiguousarray(enc)
    lib().orc_apply(psi.ctypes.data, n, enc.ctypes.data, len(enc))


def simulate_gates(gates, n_qubits: int, input=0) -> np.ndarray:
    n = max(1, int(n_qubits))
    if isinstance(input, np.ndarray):
        psi = np.ascontiguousarray(input, dtype=np.complex128).reshape(-1).copy()
    else:
        psi = np.empty(2**n, dtype=np.complex128)
        lib().orc_set_basis(psi.ctypes.data, n, int(input))
    apply_encoded(psi, n, encode_gates(gates))
    return psi


def project_norm2(psi: np.ndarray, member: np.ndarray) -> float:
    member = np.ascontiguousarray(member, dtype=np.uint8)
    t = int(member.size).bit_length() - 1
    assert member.size == 2**t
    return float(lib().orc_project_norm2(psi.ctypes.data, member.ctypes.data, t))

"""ctypes binding of ``include/b2sv.h`` -- the only way the Python host code reaches the GPU.

This is the stub a unitaria maintainer would add next to ``src/unitaria/circuit.py`` (see
INTEGRATION.md).  It fails loudly: if the shared library is missing or cannot be loaded an
``ImportError``-like :class:`EngineUnavailable` is raised; if there is no CUDA device every compute
call raises :class:`B2svError` (code ``B2SV_ENODEV``).  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2SV_LIB: load another build of the same ABI (the range-checking debug build, unitaria_b200.build --bounds)
LIB_PATH = os.environ.get("B2SV_LIB") or os.path.join(_HERE, "libunitaria_b200.so")

B2SV_C128, B2SV_C64 = 0, 1
OP_MAT1, OP_DIAG1, OP_X, OP_SWAP, OP_PHASE = 0, 1, 2, 3, 4
SEG_GATE1Q, SEG_TILE, SEG_PERM, SEG_LOW6 = 0, 1, 2, 3
CLASS_NAMES = ("gate1q", "tile", "perm", "scale", "project", "init", "swap", "perm_jit", "low6", "dist_perm")
IPC_BYTES = 320
DIST_STATE_WORDS = 8

# struct b2sv_gate (80 bytes)
GATE_DTYPE = np.dtype(
    {
        "names": ["op", "n_targets", "t0", "t1", "reserved", "ctrl_mask", "m"],
        "formats": [np.uint8, np.uint8, np.uint8, np.uint8, np.uint32, np.uint64, (np.float64, 8)],
        "offsets": [0, 1, 2, 3, 4, 8, 16],
        "itemsize": 80,
    }
)

# struct b2sv_subinstr (32 bytes)
SUB_END, SUB_ZEROS, SUB_FREE, SUB_BRANCH, SUB_REJECT = 0, 1, 2, 3, 4
SUBINSTR_DTYPE = np.dtype(
    {
        "names": ["op", "pos", "len", "pad", "next0", "next1", "weight", "add", "reserved"],
        "formats": [np.uint8, np.uint8, np.uint8, np.uint8, np.uint16, np.uint16, np.uint64, np.uint64, np.uint64],
        "offsets": [0, 1, 2, 3, 4, 6, 8, 16, 24],
        "itemsize": 32,
    }
)


class PlanOpts(ctypes.Structure):
    _fields_ = [
        ("fuse", ctypes.c_int32),
        ("tile_qubits", ctypes.c_int32),
        ("perm_runs", ctypes.c_int32),
        ("timing", ctypes.c_int32),
        ("lookahead", ctypes.c_int32),
        ("tile_io", ctypes.c_int32),
        ("jit", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]

    @classmethod
    def make(cls, fuse=True, tile_qubits=0, perm_runs=True, timing=False, lookahead=0, plain_tile_io=False, jit="auto", mux=True, tile_io=None, phases=True, low6=True, mono=True, frame=True, tail_window=False):
        o = cls()
        o.fuse = int(bool(fuse))
        o.tile_qubits = int(tile_qubits)
        o.perm_runs = int(bool(perm_runs))
        o.timing = int(bool(timing))
        o.lookahead = int(lookahead)
        o.tile_io = (1 if plain_tile_io else 0) if tile_io is None else int(tile_io)
        o.jit = {"auto": 0, "off": 1, "sync": 2}[jit]
        o.reserved = (0 if mux else 1) | (0 if phases else 2) | (0 if low6 else 4) | (0 if mono else 8) | (0 if frame else 16) | (32 if tail_window else 0)
        return o

    def with_tail_window(self):
        """Copy of these options with the "a permutation sweep follows" promise set."""
        o = type(self)()
        ctypes.memmove(ctypes.byref(o), ctypes.byref(self), ctypes.sizeof(self))
        o.reserved |= 32
        return o


class Stats(ctypes.Structure):
    _fields_ = [
        ("gates_in", ctypes.c_uint64),
        ("gates_live", ctypes.c_uint64),
        ("launches", ctypes.c_uint64),
        ("launches_by_class", ctypes.c_uint64 * 12),
        ("alg_bytes_by_class", ctypes.c_double * 12),
        ("ms_by_class", ctypes.c_double * 12),
        ("apply_ms", ctypes.c_double),
    ]

    def as_dict(self) -> dict:
        return {
            "gates_in": int(self.gates_in),
            "gates_live": int(self.gates_live),
            "launches": int(self.launches),
            "launches_by_class": {CLASS_NAMES[i]: int(self.launches_by_class[i]) for i in range(len(CLASS_NAMES))},
            "alg_bytes_by_class": {CLASS_NAMES[i]: float(self.alg_bytes_by_class[i]) for i in range(len(CLASS_NAMES))},
            "ms_by_class": {CLASS_NAMES[i]: float(self.ms_by_class[i]) for i in range(len(CLASS_NAMES))},
            "apply_ms": float(self.apply_ms),
        }


class EngineUnavailable(RuntimeError):
    """libunitaria_b200.so is missing or cannot be loaded."""


class B2svError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"b2sv error {code}: {message}")
        self.code = code


# every symbol include/b2sv.h declares: (name, restype, argtypes)
_P = ctypes.c_void_p
SYMBOLS = {
    "b2sv_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_P)]),
    "b2sv_destroy": (ctypes.c_int, [_P]),
    "b2sv_set_basis": (ctypes.c_int, [_P, ctypes.c_uint64]),
    "b2sv_set_zero": (ctypes.c_int, [_P]),
    "b2sv_upload": (ctypes.c_int, [_P, _P]),
    "b2sv_fill_random": (ctypes.c_int, [_P, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64]),
    "b2sv_download": (ctypes.c_int, [_P, _P]),
    "b2sv_apply": (ctypes.c_int, [_P, _P, ctypes.c_size_t, ctypes.POINTER(PlanOpts)]),
    "b2sv_project_norm2": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "b2sv_project_norm2_at": (
        ctypes.c_int,
        [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double)],
    ),
    "b2sv_project_gather": (
        ctypes.c_int,
        [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, _P, ctypes.c_uint64],
    ),
    "b2sv_project_gather_sparse": (
        ctypes.c_int,
        [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_uint64, ctypes.c_uint64, _P, _P, ctypes.POINTER(ctypes.c_uint64)],
    ),
    "b2sv_project_gather_async": (
        ctypes.c_int,
        [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, _P, ctypes.c_uint64, ctypes.c_int],
    ),
    "b2sv_gather_wait": (ctypes.c_int, [_P, ctypes.c_int]),
    "b2sv_project_gather_device": (
        ctypes.c_int,
        [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.POINTER(_P), ctypes.c_uint64],
    ),
    "b2sv_upload_device": (ctypes.c_int, [_P, _P]),
    "b2sv_embed": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_uint64]),
    "b2sv_enumerate_basis": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_uint64]),
    "b2sv_device_ptr": (ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(ctypes.c_uint64)]),
    "b2sv_sync": (ctypes.c_int, [_P]),
    "b2sv_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_P)]),
    "b2sv_host_free": (ctypes.c_int, [_P]),
    "b2sv_mark": (ctypes.c_int, [_P, ctypes.c_int]),
    "b2sv_elapsed_ms": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "b2sv_swap_pack": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "b2sv_swap_unpack": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "b2sv_swap_pack_range": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P, ctypes.c_uint64, ctypes.c_uint64]),
    "b2sv_swap_unpack_range": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P, ctypes.c_uint64, ctypes.c_uint64]),
    "b2sv_ipc_export": (ctypes.c_int, [_P, _P]),
    "b2sv_ipc_attach": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "b2sv_ipc_detach": (ctypes.c_int, [_P]),
    "b2sv_dist_ready": (ctypes.c_int, [_P, _P]),
    "b2sv_dist_perm": (ctypes.c_int, [_P, _P, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, _P, ctypes.POINTER(PlanOpts)]),
    "b2sv_dist_done": (ctypes.c_int, [_P]),
    "b2sv_dist_source": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_size_t, _P, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "b2sv_plan_describe": (
        ctypes.c_int,
        [ctypes.c_int, ctypes.c_int, _P, ctypes.c_size_t, ctypes.POINTER(PlanOpts), _P, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)],
    ),
    "b2sv_perm_jit_compile_check": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _P, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "b2sv_stats": (ctypes.c_int, [_P, ctypes.POINTER(Stats)]),
    "b2sv_stats_reset": (ctypes.c_int, [_P]),
    "b2sv_abi_version": (ctypes.c_int, []),
    "b2sv_device_count": (ctypes.c_int, []),
    "b2sv_last_error": (ctypes.c_char_p, []),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the C-ABI library (built by ``unitaria_b200.build`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            f"{LIB_PATH} not found: build it with `python -m unitaria_b200.build` (needs nvcc). "
            "unitaria_b200 has no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the host
        raise EngineUnavailable(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b2sv_last_error()
        raise B2svError(rc, msg.decode() if msg else "unknown error")


def device_count() -> int:
    return int(load().b2sv_device_count())


def plan_describe(n_qubits: int, dtype: int, gates: np.ndarray, opts: PlanOpts | None = None, with_common: bool = False, with_relabel: bool = False):
    """Planner introspection (host only, needs no GPU): list of (class, tile_mask, [gate indices]) and
    the folded global scalar.  ``with_relabel``: also the {gate id: (op, t0, t1, ctrl_mask)} table of gates the
    SWAP frame executes on renamed qubits (ids >= len(gates) are synthetic layout-restoring SWAPs)."""
    lib = load()
    gates = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
    need = ctypes.c_size_t(0)
    optp = ctypes.byref(opts) if opts is not None else None
    check(lib.b2sv_plan_describe(n_qubits, dtype, gates.ctypes.data, len(gates), optp, None, 0, ctypes.byref(need)))
    words = np.zeros(need.value, dtype=np.int64)
    check(lib.b2sv_plan_describe(n_qubits, dtype, gates.ctypes.data, len(gates), optp, words.ctypes.data, len(words), ctypes.byref(need)))
    segs = []
    scalar = 1.0 + 0.0j
    i = 0
    common = []
    relabel = {}
    while i < len(words):
        cls, cnt, mask, com, n_ops = (int(words[i + k]) for k in range(5))
        if cls == -2:
            for r in range(cnt):
                gid, op, t0, t1, ctrl = (int(words[i + 5 + 5 * r + k]) for k in range(5))
                relabel[gid] = (op, t0, t1, ctrl & 0xFFFFFFFFFFFFFFFF)
            i += 5 + 5 * cnt
            continue
        if cls == -1:
            re = np.array([words[i + 5]], dtype=np.int64).view(np.float64)[0]
            im = np.array([words[i + 6]], dtype=np.int64).view(np.float64)[0]
            scalar = complex(re, im)
            i += 7
            continue
        segs.append((cls, mask & 0xFFFFFFFFFFFFFFFF, [int(w) for w in words[i + 5 : i + 5 + cnt]]))
        common.append((com & 0xFFFFFFFFFFFFFFFF, n_ops))
        i += 5 + cnt
    if with_relabel:
        return segs, scalar, common, relabel
    if with_common:
        return segs, scalar, common
    return segs, scalar

"""Host-side handle on one device-resident statevector (thin wrapper over the C ABI)."""

from __future__ import annotations

import ctypes
import weakref

import numpy as np

from . import cabi
from .cabi import B2SV_C64, B2SV_C128, GATE_DTYPE, PlanOpts, Stats
from .lowering import lower_gates
from .subspace_prog import SubspaceProgram

_DTYPES = {
    "complex128": B2SV_C128,
    "c128": B2SV_C128,
    np.complex128: B2SV_C128,
    "complex64": B2SV_C64,
    "c64": B2SV_C64,
    np.complex64: B2SV_C64,
}


class _PinnedPool:
    """Recycles page-locked host buffers for results that cross PCIe.

    ``cudaHostAlloc`` is slow (it pins pages), so buffers are handed out as numpy arrays and come
    back to the pool when the last array (or view) on them is garbage-collected.  A caller that
    keeps a result alive simply keeps its buffer; nothing is ever overwritten behind its back.
    """

    MIN_BYTES = 1 << 20
    MAX_CACHED = 6 << 30

    def __init__(self):
        self.free: dict = {}
        self.cached = 0

    def array(self, count: int, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        nbytes = int(count) * dtype.itemsize
        if nbytes < self.MIN_BYTES:
            return np.empty(count, dtype=dtype)
        lib = cabi.load()
        bucket = self.free.get(nbytes)
        if bucket:
            ptr = bucket.pop()
            self.cached -= nbytes
        else:
            p = ctypes.c_void_p()
            if lib.b2sv_host_alloc(nbytes, ctypes.byref(p)) != 0:
                return np.empty(count, dtype=dtype)  # pinning failed: pageable memory still works
            ptr = p.value
        buf = (ctypes.c_char * nbytes).from_address(ptr)
        weakref.finalize(buf, self._release, ptr, nbytes)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def _release(self, ptr: int, nbytes: int) -> None:
        if self.cached + nbytes <= self.MAX_CACHED:
            self.free.setdefault(nbytes, []).append(ptr)
            self.cached += nbytes
        else:
            try:
                cabi.load().b2sv_host_free(ctypes.c_void_p(ptr))
            except Exception:
                pass

    def clear(self) -> None:
        lib = cabi.load()
        for bucket in self.free.values():
            for ptr in bucket:
                lib.b2sv_host_free(ctypes.c_void_p(ptr))
        self.free.clear()
        self.cached = 0


PINNED = _PinnedPool()


def _dtype_code(dtype) -> int:
    if dtype in (B2SV_C128, B2SV_C64) and isinstance(dtype, int):
        return dtype
    try:
        return _DTYPES[dtype]
    except (KeyError, TypeError):
        return _DTYPES[np.dtype(dtype).name]


class StateVector:
    """``2**n_qubits`` amplitudes on one B200, LSB order (index bit q = qubit q).

    ``n_qubits`` follows ``Circuit.simulate``'s register rule ``max(1, n_qubits)``
    (/root/reference/src/unitaria/circuit.py:65).  Host arrays in and out are complex128 whatever
    the device storage dtype is.
    """

    def __init__(self, n_qubits: int, dtype="complex128", device: int = 0):
        self._lib = cabi.load()
        self.n_qubits = max(1, int(n_qubits))
        self.dtype_code = _dtype_code(dtype)
        self.device = int(device)
        handle = ctypes.c_void_p()
        cabi.check(self._lib.b2sv_create(self.n_qubits, self.dtype_code, self.device, ctypes.byref(handle)))
        self._h = handle

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.b2sv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def size(self) -> int:
        return 1 << self.n_qubits

    @property
    def amp_bytes(self) -> int:
        return 16 if self.dtype_code == B2SV_C128 else 8

    # -- state I/O ------------------------------------------------------------------------------
    def set_basis(self, index: int) -> None:
        cabi.check(self._lib.b2sv_set_basis(self._h, int(index)))

    def set_zero(self) -> None:
        cabi.check(self._lib.b2sv_set_zero(self._h))

    def upload(self, amplitudes: np.ndarray) -> None:
        a = np.ascontiguousarray(amplitudes, dtype=np.complex128).reshape(-1)
        if a.size != self.size:
            raise ValueError(f"expected {self.size} amplitudes, got {a.size}")
        cabi.check(self._lib.b2sv_upload(self._h, a.ctypes.data))

    def download(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = PINNED.array(self.size, np.complex128)
        assert out.dtype == np.complex128 and out.flags.c_contiguous and out.size == self.size
        cabi.check(self._lib.b2sv_download(self._h, out.ctypes.data))
        return out

    def load(self, input) -> None:
        """``int`` -> basis state, ndarray -> dense vector (circuit.py:62-65)."""
        if isinstance(input, np.ndarray):
            self.upload(input)
        elif isinstance(input, (int, np.integer)):
            self.set_basis(int(input))
        else:
            raise TypeError(f"input must be an int or an ndarray, not {type(input).__name__}")

    # -- gates ----------------------------------------------------------------------------------
    def apply_lowered(self, lowered: np.ndarray, opts: PlanOpts | None = None) -> None:
        lowered = np.ascontiguousarray(lowered, dtype=GATE_DTYPE)
        optp = ctypes.byref(opts) if opts is not None else None
        cabi.check(self._lib.b2sv_apply(self._h, lowered.ctypes.data, len(lowered), optp))

    def apply(self, gates, opts: PlanOpts | None = None) -> None:
        """Apply tequila-like gates (lowered on the fly)."""
        self.apply_lowered(lower_gates(gates, self.n_qubits), opts)

    # -- projection -----------------------------------------------------------------------------
    @staticmethod
    def _program(subspace) -> SubspaceProgram:
        return subspace if isinstance(subspace, SubspaceProgram) else SubspaceProgram(subspace)

    def project_norm2(self, subspace) -> float:
        """sum |psi_i|^2 over the subspace on the low qubits, all higher qubits 0 (node.py:379,402)."""
        p = self._program(subspace)
        out = ctypes.c_double(0.0)
        cabi.check(
            self._lib.b2sv_project_norm2(self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, ctypes.byref(out))
        )
        return float(out.value)

    def project_norm2_at(self, subspace, index_base: int) -> float:
        """Shard form of :meth:`project_norm2`: this register holds indices ``index_base | i`` of a wider one."""
        p = self._program(subspace)
        out = ctypes.c_double(0.0)
        cabi.check(
            self._lib.b2sv_project_norm2_at(
                self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, int(index_base), ctypes.byref(out)
            )
        )
        return float(out.value)

    def project_norm2_shard(self, shard_prog) -> float:
        """Norm over a per-rank specialised membership program (subspace_prog.ShardProgram)."""
        if shard_prog.rejects_all:
            return 0.0
        out = ctypes.c_double(0.0)
        cabi.check(
            self._lib.b2sv_project_norm2_at(
                self._h, shard_prog.instrs.ctypes.data, len(shard_prog.instrs), shard_prog.entry, self.n_qubits, 0, ctypes.byref(out)
            )
        )
        return float(out.value)

    def project(self, subspace, scale: complex = 1.0) -> np.ndarray:
        """``scale * Subspace.project(psi)`` (subspace.py:331-335) gathered on the device."""
        p = self._program(subspace)
        out = PINNED.array(p.dimension, np.complex128)
        s = complex(scale)
        cabi.check(
            self._lib.b2sv_project_gather(
                self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, s.real, s.imag, out.ctypes.data, p.dimension
            )
        )
        return out

    def embed(self, subspace, vector: np.ndarray) -> None:
        """psi := sum_r vector[r] |basis_r>: the inverse of :meth:`project` and the device form of the
        embedding in ``BlockEncoding.compute`` (nodes/block_encoding.py:49-52) -- only ``dimension``
        amplitudes are uploaded, the rest of the register is zero-filled on the device."""
        p = self._program(subspace)
        v = np.ascontiguousarray(vector, dtype=np.complex128).reshape(-1)
        if v.size != p.dimension:
            raise ValueError(f"vector has {v.size} entries, the subspace has dimension {p.dimension}")
        cabi.check(self._lib.b2sv_embed(self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, v.ctypes.data, p.dimension))

    def enumerate_basis(self, subspace) -> np.ndarray:
        """``Subspace.enumerate_basis`` (subspace.py:325-329) with int64 indices."""
        p = self._program(subspace)
        out = np.empty(p.dimension, dtype=np.int64)
        cabi.check(
            self._lib.b2sv_enumerate_basis(self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, out.ctypes.data, p.dimension)
        )
        return out

    # -- hand-off / bookkeeping -------------------------------------------------------------------
    def device_ptr(self) -> tuple[int, int]:
        ptr = ctypes.c_void_p()
        nbytes = ctypes.c_uint64()
        cabi.check(self._lib.b2sv_device_ptr(self._h, ctypes.byref(ptr), ctypes.byref(nbytes)))
        return int(ptr.value), int(nbytes.value)

    def sync(self) -> None:
        cabi.check(self._lib.b2sv_sync(self._h))

    def mark(self, slot: int) -> None:
        """Record a CUDA event on the engine's stream (stopwatch slot 0..7)."""
        cabi.check(self._lib.b2sv_mark(self._h, int(slot)))

    def elapsed_ms(self, slot_a: int, slot_b: int) -> float:
        ms = ctypes.c_double(0.0)
        cabi.check(self._lib.b2sv_elapsed_ms(self._h, int(slot_a), int(slot_b), ctypes.byref(ms)))
        return float(ms.value)

    def stats(self) -> dict:
        st = Stats()
        cabi.check(self._lib.b2sv_stats(self._h, ctypes.byref(st)))
        return st.as_dict()

    def stats_reset(self) -> None:
        cabi.check(self._lib.b2sv_stats_reset(self._h))

    def swap_pack(self, qubit: int, keep_bit: int, staging_ptr: int, first: int = 0, count: int | None = None) -> None:
        count = (self.size >> 1) - first if count is None else count
        cabi.check(self._lib.b2sv_swap_pack_range(self._h, int(qubit), int(keep_bit), ctypes.c_void_p(staging_ptr), int(first), int(count)))

    def swap_unpack(self, qubit: int, keep_bit: int, staging_ptr: int, first: int = 0, count: int | None = None) -> None:
        count = (self.size >> 1) - first if count is None else count
        cabi.check(self._lib.b2sv_swap_unpack_range(self._h, int(qubit), int(keep_bit), ctypes.c_void_p(staging_ptr), int(first), int(count)))

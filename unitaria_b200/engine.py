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


def random_state_hash(n_qubits: int, seed: int = 0, support_bits: int | None = None, index_base: int = 0) -> np.ndarray:
    """Host restatement of ``b2sv_fill_random`` (splitmix64 counter hash), for parity checks at small sizes."""
    bits = n_qubits if support_bits is None else support_bits
    k = np.arange(2**n_qubits, dtype=np.uint64) + np.uint64(index_base)

    def unit(x):
        with np.errstate(over="ignore"):
            x = x + np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15)
            x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            x = x ^ (x >> np.uint64(31))
        return (x >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0

    scale = np.sqrt(1.5 / 2.0**bits)
    out = scale * (unit(np.uint64(2) * k) + 1j * unit(np.uint64(2) * k + np.uint64(1)))
    out[k >= np.uint64(2**bits)] = 0
    return out


class _DeviceArray:
    """``__cuda_array_interface__`` view of a device buffer owned by a StateVector (tensor hand-off only)."""

    def __init__(self, ptr: int, count: int, owner):
        self.owner = owner  # keeps the handle (and with it the buffer) alive
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<c16", "data": (ptr, False), "version": 2, "strides": None}


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

    def fill_random(self, seed: int = 0, support_bits: int | None = None, index_base: int = 0) -> None:
        """Seeded dense vector generated on the device (``random_state_hash`` below is its host restatement)."""
        bits = self.n_qubits if support_bits is None else int(support_bits)
        cabi.check(self._lib.b2sv_fill_random(self._h, int(seed), bits, int(index_base)))

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

    SPARSE_MIN_DIM = 1 << 16   # below this the dense copy is a few microseconds anyway
    SPARSE_MAX_PAIRS = 1 << 18

    def project_sparse(self, subspace, scale: complex = 1.0):
        """The non-zero entries of :meth:`project` as ``(ranks, values)`` (no particular order), or ``None`` when
        there are too many of them for the compact form (then :meth:`project` is the way).  8 + 24 bytes per
        non-zero cross PCIe instead of 16 bytes per subspace dimension."""
        p = self._program(subspace)
        cap = int(min(max(p.dimension >> 6, 1024), self.SPARSE_MAX_PAIRS))
        bufs = getattr(self, "_sparse_bufs", None)
        if bufs is None or len(bufs[0]) < cap:
            bufs = self._sparse_bufs = (np.empty(cap, dtype=np.uint64), np.empty(cap, dtype=np.complex128))
        idx, val = bufs
        nnz = ctypes.c_uint64(0)
        s = complex(scale)
        cabi.check(
            self._lib.b2sv_project_gather_sparse(
                self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, s.real, s.imag, p.dimension, cap,
                idx.ctypes.data, val.ctypes.data, ctypes.byref(nnz),
            )
        )
        n = int(nnz.value)
        if n >= cap:
            return None
        return idx[:n].copy(), val[:n].copy()

    def project(self, subspace, scale: complex = 1.0) -> np.ndarray:
        """``scale * Subspace.project(psi)`` (subspace.py:331-335) gathered on the device.  Large results are tried in
        compact form first (a block-encoding column is mostly exact zeros): the dense array is then assembled on the
        host from the non-zeros.  ``last_d2h_bytes`` says what crossed PCIe."""
        p = self._program(subspace)
        skip = getattr(self, "_sparse_skip", 0)
        if p.dimension >= self.SPARSE_MIN_DIM and skip == 0:
            pairs = self.project_sparse(p, scale)
            if pairs is not None:
                out = np.zeros(p.dimension, dtype=np.complex128)  # calloc: untouched pages cost nothing
                out[pairs[0]] = pairs[1]
                self.last_d2h_bytes = 8 + 24 * len(pairs[0])
                return out
            self._sparse_skip = 8  # dense result: the next few of this handle probably are as well
        elif skip:
            self._sparse_skip = skip - 1
        out = PINNED.array(p.dimension, np.complex128)
        self.last_d2h_bytes = 16 * p.dimension
        s = complex(scale)
        cabi.check(
            self._lib.b2sv_project_gather(
                self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, s.real, s.imag, out.ctypes.data, p.dimension
            )
        )
        return out

    def project_device(self, subspace, scale: complex = 1.0):
        """:meth:`project` with the result left on the GPU: a torch complex128 tensor VIEWING the handle's result
        buffer (valid until the next projection on this handle -- ``.clone()`` it to keep it)."""
        import torch

        p = self._program(subspace)
        ptr = ctypes.c_void_p()
        s = complex(scale)
        cabi.check(
            self._lib.b2sv_project_gather_device(
                self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, s.real, s.imag, ctypes.byref(ptr), p.dimension
            )
        )
        if p.dimension == 0:
            return torch.empty(0, dtype=torch.complex128, device=f"cuda:{self.device}")
        view = _DeviceArray(int(ptr.value), p.dimension, self)
        return torch.as_tensor(view, device=f"cuda:{self.device}")

    def upload_device(self, tensor) -> None:
        """Load the state from a torch CUDA tensor (complex128, 2^n amplitudes, same GPU): no host round trip."""
        import torch

        t = tensor.contiguous().reshape(-1)
        if t.dtype != torch.complex128 or t.numel() != self.size or not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"expected a complex128 CUDA tensor with {self.size} amplitudes on cuda:{self.device}")
        torch.cuda.current_stream(t.device).synchronize()
        cabi.check(self._lib.b2sv_upload_device(self._h, ctypes.c_void_p(t.data_ptr())))

    def project_async(self, subspace, scale: complex, out: np.ndarray, slot: int) -> None:
        """:meth:`project` without waiting: the result lands in ``out`` once :meth:`gather_wait` (same slot) returns."""
        p = self._program(subspace)
        assert out.dtype == np.complex128 and out.flags.c_contiguous and out.size == p.dimension
        s = complex(scale)
        cabi.check(
            self._lib.b2sv_project_gather_async(
                self._h, p.instrs.ctypes.data, len(p.instrs), p.entry, p.total_qubits, s.real, s.imag, out.ctypes.data, p.dimension, int(slot)
            )
        )

    def gather_wait(self, slot: int) -> None:
        cabi.check(self._lib.b2sv_gather_wait(self._h, int(slot)))

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

    # -- distributed execution (unitaria_b200/distributed.py) ---------------------------------------
    def ipc_export(self) -> bytes:
        """Opaque blob (state buffers + hand-shake events) that peers attach with :meth:`ipc_attach`."""
        blob = ctypes.create_string_buffer(cabi.IPC_BYTES)
        cabi.check(self._lib.b2sv_ipc_export(self._h, blob))
        return blob.raw

    def ipc_attach(self, rank: int, world: int, blobs) -> None:
        raw = b"".join(blobs)
        if len(raw) != world * cabi.IPC_BYTES:
            raise ValueError("one IPC blob per rank is needed")
        cabi.check(self._lib.b2sv_ipc_attach(self._h, int(rank), int(world), raw))

    def ipc_detach(self) -> None:
        cabi.check(self._lib.b2sv_ipc_detach(self._h))

    def dist_ready(self) -> np.ndarray:
        state = np.zeros(cabi.DIST_STATE_WORDS, dtype=np.uint64)
        cabi.check(self._lib.b2sv_dist_ready(self._h, state.ctypes.data))
        return state

    def dist_perm(self, lowered: np.ndarray, n_total: int, flip_in: int, flip_out: int, states: np.ndarray, opts: PlanOpts | None = None) -> None:
        lowered = np.ascontiguousarray(lowered, dtype=GATE_DTYPE)
        states = np.ascontiguousarray(states, dtype=np.uint64)
        optp = ctypes.byref(opts) if opts is not None else None
        cabi.check(self._lib.b2sv_dist_perm(self._h, lowered.ctypes.data, len(lowered), int(n_total), int(flip_in), int(flip_out), states.ctypes.data, optp))

    def dist_done(self) -> None:
        cabi.check(self._lib.b2sv_dist_done(self._h))

    def swap_pack(self, qubit: int, keep_bit: int, staging_ptr: int, first: int = 0, count: int | None = None) -> None:
        count = (self.size >> 1) - first if count is None else count
        cabi.check(self._lib.b2sv_swap_pack_range(self._h, int(qubit), int(keep_bit), ctypes.c_void_p(staging_ptr), int(first), int(count)))

    def swap_unpack(self, qubit: int, keep_bit: int, staging_ptr: int, first: int = 0, count: int | None = None) -> None:
        count = (self.size >> 1) - first if count is None else count
        cabi.check(self._lib.b2sv_swap_unpack_range(self._h, int(qubit), int(keep_bit), ctypes.c_void_p(staging_ptr), int(first), int(count)))

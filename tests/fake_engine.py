"""TEST DOUBLE for the CUDA engine: executes the C ABI's data formats (lowered ``b2sv_gate`` arrays and
compiled subspace programs) with numpy.  It exists so that the host logic -- lowering, subspace
compilation, the adapter's patched ``Circuit.simulate`` / ``Node.simulate`` -- can be run against the
reference's own test-suite on a machine without a GPU.  It lives under tests/ and is never imported by
the product."""

from __future__ import annotations

import numpy as np

from unitaria_b200 import cabi
from unitaria_b200.subspace_prog import SubspaceProgram


def apply_lowered(psi: np.ndarray, n: int, lowered: np.ndarray) -> None:
    view = psi.reshape([2] * n)

    def sl(ctrl_mask, fixed=None):
        idx = [slice(None)] * n
        for q in range(n):
            if (ctrl_mask >> q) & 1:
                idx[n - 1 - q] = 1
        for q, v in (fixed or {}).items():
            idx[n - 1 - q] = v
        return tuple(idx)

    for g in lowered:
        op, t0, t1, cm, m = int(g["op"]), int(g["t0"]), int(g["t1"]), int(g["ctrl_mask"]), g["m"]
        c = lambda i: complex(m[2 * i], m[2 * i + 1])  # noqa: E731
        if op == cabi.OP_PHASE:
            view[sl(cm)] *= c(0)
        elif op == cabi.OP_X:
            s0, s1 = sl(cm, {t0: 0}), sl(cm, {t0: 1})
            tmp = view[s0].copy()
            view[s0] = view[s1]
            view[s1] = tmp
        elif op == cabi.OP_SWAP:
            s01, s10 = sl(cm, {t0: 1, t1: 0}), sl(cm, {t0: 0, t1: 1})
            tmp = view[s01].copy()
            view[s01] = view[s10]
            view[s10] = tmp
        elif op == cabi.OP_DIAG1:
            s0, s1 = sl(cm, {t0: 0}), sl(cm, {t0: 1})
            view[s0] *= c(0)
            view[s1] *= c(1)
        else:
            s0, s1 = sl(cm, {t0: 0}), sl(cm, {t0: 1})
            a0, a1 = view[s0].copy(), view[s1].copy()
            view[s0] = c(0) * a0 + c(1) * a1
            view[s1] = c(2) * a0 + c(3) * a1


def program_members(prog: SubspaceProgram):
    """(member indices in rank order) by running the compiled decision program on every index."""
    ins = prog.instrs
    out = {}
    for i in range(2**prog.total_qubits):
        pc, rank, ok = prog.entry, 0, None
        for _ in range(len(ins) + 1):
            op, pos, ln = int(ins["op"][pc]), int(ins["pos"][pc]), int(ins["len"][pc])
            field = (i >> pos) & ((1 << ln) - 1)
            if op == cabi.SUB_END:
                ok = True
                break
            if op == cabi.SUB_REJECT or (op == cabi.SUB_ZEROS and field):
                ok = False
                break
            if op == cabi.SUB_ZEROS:
                pc = int(ins["next0"][pc])
            elif op == cabi.SUB_FREE:
                rank += field * int(ins["weight"][pc])
                pc = int(ins["next0"][pc])
            elif (i >> pos) & 1:
                rank += int(ins["add"][pc])
                pc = int(ins["next1"][pc])
            else:
                pc = int(ins["next0"][pc])
        assert ok is not None
        if ok:
            out[rank] = i
    assert sorted(out) == list(range(len(out)))
    return np.array([out[r] for r in range(len(out))], dtype=np.int64)


class FakeStateVector:
    """Same surface as unitaria_b200.engine.StateVector (the part the adapter uses)."""

    _members_cache: dict = {}

    def __init__(self, n_qubits, dtype="complex128", device=0):
        self.n_qubits = max(1, int(n_qubits))
        self.size = 1 << self.n_qubits
        self.psi = np.zeros(self.size, dtype=np.complex128)
        self.psi[0] = 1

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def project_async(self, subspace, scale, out, slot):
        out[:] = self.project(subspace, scale)

    def gather_wait(self, slot):
        pass

    def set_basis(self, index):
        self.psi[:] = 0
        self.psi[int(index)] = 1

    def upload(self, a):
        self.psi[:] = np.asarray(a, dtype=np.complex128).reshape(-1)

    def download(self, out=None):
        return self.psi.copy()

    def apply_lowered(self, lowered, opts=None):
        apply_lowered(self.psi, self.n_qubits, lowered)

    def _members(self, prog):
        key = prog.instrs.tobytes() + bytes([prog.entry & 0xFF, prog.entry >> 8])
        hit = self._members_cache.get(key)
        if hit is None:
            hit = program_members(prog)
            self._members_cache[key] = hit
        return hit

    def enumerate_basis(self, subspace):
        prog = subspace if isinstance(subspace, SubspaceProgram) else SubspaceProgram(subspace)
        return self._members(prog)

    def embed(self, subspace, vector):
        prog = subspace if isinstance(subspace, SubspaceProgram) else SubspaceProgram(subspace)
        self.psi[:] = 0
        self.psi[self._members(prog)] = np.asarray(vector, dtype=np.complex128).reshape(-1)

    def project(self, subspace, scale=1.0):
        prog = subspace if isinstance(subspace, SubspaceProgram) else SubspaceProgram(subspace)
        return complex(scale) * self.psi[self._members(prog)]

    SPARSE_MIN_DIM = 1 << 16

    def project_sparse(self, subspace, scale=1.0):
        dense = self.project(subspace, scale)
        nz = np.flatnonzero(dense)
        if len(nz) >= max(len(dense) >> 6, 1024):
            return None
        return nz.astype(np.uint64), dense[nz]

    def project_norm2(self, subspace):
        prog = subspace if isinstance(subspace, SubspaceProgram) else SubspaceProgram(subspace)
        return float(np.sum(np.abs(self.psi[self._members(prog)]) ** 2))

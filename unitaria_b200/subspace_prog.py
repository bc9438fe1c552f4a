"""Subspace -> stack-free device program (``b2sv_subinstr[]``), SURVEY.md App. C.

A unitaria ``Subspace`` (/root/reference/src/unitaria/subspace.py) is a tensor product of factors,
least-significant first (:117-125): ``ZeroQubitSubspace`` (one qubit fixed to 0) or
``ControlledSubspace(case_zero, case_one)`` whose top qubit picks which sub-subspace constrains the
lower qubits (:523-560).  ``test_basis`` (:297-323) walks those factors; the device runs the same
walk as a compiled decision program, plus the mixed-radix rank that reproduces ``enumerate_basis``
order (:325-329) without enumerating.

Accepted inputs: a unitaria ``Subspace`` (duck-typed via ``tensor_factors`` / ``case_zero`` /
``case_one``), a string such as ``"0##"`` (most significant qubit first, subspace.py:114-125) or the
neutral nested spec used by the fixtures: a list of factors, each ``0`` or ``[spec0, spec1]``.
"""

from __future__ import annotations

import numpy as np

from .cabi import SUB_BRANCH, SUB_END, SUB_FREE, SUB_REJECT, SUB_ZEROS, SUBINSTR_DTYPE


def to_spec(subspace) -> list:
    if isinstance(subspace, str):
        # leftmost char = most significant qubit; "#" is a free qubit, "0" a zero qubit
        spec = []
        for ch in reversed(subspace):
            if ch == "0":
                spec.append(0)
            elif ch == "#":
                spec.append([[], []])
            else:
                raise ValueError(f"bad subspace string {subspace!r}")
        return spec
    if isinstance(subspace, (list, tuple)):
        return [0 if f == 0 else [to_spec(f[0]), to_spec(f[1])] for f in subspace]
    factors = getattr(subspace, "tensor_factors", None)
    if factors is None:
        raise TypeError(f"cannot interpret {type(subspace).__name__} as a Subspace")
    spec = []
    for f in factors:
        kind = type(f).__name__
        if kind == "ZeroQubitSubspace":
            spec.append(0)
        elif kind == "ControlledSubspace":
            spec.append([to_spec(f.case_zero), to_spec(f.case_one)])
        else:
            raise NotImplementedError(f"subspace factor {kind}")
    return spec


def spec_total_qubits(spec) -> int:
    return sum(1 if f == 0 else 1 + spec_total_qubits(f[0]) for f in spec)


def spec_dimension(spec) -> int:
    d = 1
    for f in spec:
        if f != 0:
            d *= spec_dimension(f[0]) + spec_dimension(f[1])
    return d


def spec_member(spec, rank: int) -> int:
    """The ``rank``-th member of the subspace in ``enumerate_basis`` order (ascending index,
    subspace.py:325-329), without enumerating: the rank is a mixed-radix number over the factors, factor 0
    least significant; the inverse of the rank the device programs compute."""
    index, pos = 0, 0
    for f in spec:
        if f == 0:
            pos += 1
            continue
        d0, d1 = spec_dimension(f[0]), spec_dimension(f[1])
        rank, r = divmod(rank, d0 + d1)
        width = spec_total_qubits(f[0])
        if r < d0:
            index |= spec_member(f[0], r) << pos
        else:
            index |= (spec_member(f[1], r - d0) | (1 << width)) << pos
        pos += width + 1
    if rank:
        raise IndexError("rank outside the subspace dimension")
    return index


class SubspaceProgram:
    """Compiled subspace: ``instrs`` (SUBINSTR_DTYPE), ``entry``, ``total_qubits``, ``dimension``.

    Because a Subspace is a tree, every factor has a static absolute bit position and a static
    mixed-radix weight (product of the dimensions of all lower factors), so the
    walk of ``test_basis`` compiles to a stack-free decision program (include/b2sv.h).
    """

    def __init__(self, subspace):
        self.spec = to_spec(subspace)
        self.total_qubits = spec_total_qubits(self.spec)
        self.dimension = spec_dimension(self.spec)
        if self.total_qubits > 64:
            raise ValueError("subspaces wider than 64 qubits are not supported")
        # "0..0#..#" (free qubits below, zero qubits above): the members are exactly 0 .. dimension-1, so
        # enumerate_basis() is arange(dimension) and projecting a length-dimension vector again is the identity
        free = [f != 0 and f == [[], []] for f in self.spec]
        n_free = sum(free)
        self.is_index_window = all(free[:n_free]) and all(f == 0 for f in self.spec[n_free:])
        self._ins: list = [dict(op=SUB_END, pos=0, len=0, next0=0, next1=0, weight=0, add=0)]
        self.entry = self._emit_list(self.spec, 0, 1, 0)
        if len(self._ins) > 65535:
            raise ValueError("subspace program too long")
        self.instrs = np.zeros(len(self._ins), dtype=SUBINSTR_DTYPE)
        for i, r in enumerate(self._ins):
            for k, v in r.items():
                self.instrs[k][i] = v
        del self._ins

    def _push(self, **fields) -> int:
        self._ins.append(fields)
        return len(self._ins) - 1

    def _emit_list(self, spec, pos: int, weight: int, cont: int) -> int:
        """Emit the program of the factor list ``spec`` whose lowest qubit sits at absolute bit ``pos``
        and whose rank is scaled by ``weight``; when the list is done control goes to ``cont``.
        Returns the entry instruction.  Built back to front so every jump target already exists."""
        positions, weights = [], []
        p, w = pos, weight
        for f in spec:
            positions.append(p)
            weights.append(w)
            if f == 0:
                p += 1
            else:
                if spec_total_qubits(f[0]) != spec_total_qubits(f[1]):
                    raise ValueError("ControlledSubspace cases span different qubit counts")
                p += 1 + spec_total_qubits(f[0])
                w *= spec_dimension(f[0]) + spec_dimension(f[1])
        nxt = cont
        for j in range(len(spec) - 1, -1, -1):
            f, fp, fw = spec[j], positions[j], weights[j]
            if fw >= 2**64:
                raise ValueError("subspace dimension exceeds 64 bits")
            head = self._ins[nxt]
            if f == 0:
                if head["op"] == SUB_ZEROS and head["pos"] == fp + 1 and nxt != cont:
                    head["pos"], head["len"] = fp, head["len"] + 1  # extend the run downwards
                else:
                    nxt = self._push(op=SUB_ZEROS, pos=fp, len=1, next0=nxt, next1=0, weight=0, add=0)
            elif len(f[0]) == 0 and len(f[1]) == 0:  # free qubit
                if head["op"] == SUB_FREE and head["pos"] == fp + 1 and head["weight"] == 2 * fw and nxt != cont:
                    head["pos"], head["len"], head["weight"] = fp, head["len"] + 1, fw
                else:
                    nxt = self._push(op=SUB_FREE, pos=fp, len=1, next0=nxt, next1=0, weight=fw, add=0)
            else:
                nb = spec_total_qubits(f[0])
                e0 = self._emit_list(f[0], fp, fw, nxt)
                e1 = self._emit_list(f[1], fp, fw, nxt)
                nxt = self._push(op=SUB_BRANCH, pos=fp + nb, len=1, next0=e0, next1=e1, weight=0, add=spec_dimension(f[0]) * fw)
        return nxt

    # kept for callers that want the flat arrays under their C names
    @property
    def nodes(self):
        return self.instrs


class ShardProgram:
    """Membership program of a subspace for ONE shard of a sharded register (norm only: no ranks).

    The logical index bits of the subspace program are relabelled to where the qubits physically
    sit (``position[q]``); bits that live in the rank number (``position >= n_local``) are known on the
    host -- ``rank_value(position)`` gives the logical value on this rank -- so their instructions are
    resolved here: a ZEROS on a 1 becomes REJECT, a BRANCH becomes a jump.  The device evaluates what
    is left on the local index alone, so the projection needs no global<->local swaps.
    """

    def __init__(self, prog: SubspaceProgram, position, n_local: int, rank_value, n_total: int | None = None):
        src = prog.instrs
        out = []

        def emit(**f):
            out.append(f)
            return len(out) - 1

        # every source instruction becomes a chain of single-bit instructions; first pass allocates the
        # head index of each chain so that jumps can be resolved
        heads, chains = {}, {}
        for i, ins in enumerate(src):
            op, pos, ln = int(ins["op"]), int(ins["pos"]), int(ins["len"])
            if op in (SUB_END, SUB_REJECT):
                chains[i] = [(op, 0)]
            elif op == SUB_BRANCH:
                chains[i] = [(op, pos)]
            elif op == SUB_ZEROS:
                chains[i] = [(op, pos + b) for b in range(ln)]
            else:  # FREE only matters for the rank of a member: a no-op for the norm
                chains[i] = [(SUB_FREE, -1)]
        cursor = 0
        for i in range(len(src)):
            heads[i] = cursor
            cursor += len(chains[i])
        for i, ins in enumerate(src):
            n0, n1 = heads[int(ins["next0"])], heads[int(ins["next1"])]
            chain = chains[i]
            for j, (op, bit) in enumerate(chain):
                nxt = heads[i] + j + 1 if j + 1 < len(chain) else n0
                if op in (SUB_END, SUB_REJECT):
                    emit(op=op, pos=0, len=0, next0=0, next1=0, weight=0, add=0)
                elif op == SUB_FREE:
                    emit(op=SUB_FREE, pos=0, len=1, next0=nxt, next1=0, weight=0, add=0)
                elif op == SUB_ZEROS:
                    p = position[bit]
                    if p < n_local:
                        emit(op=SUB_ZEROS, pos=p, len=1, next0=nxt, next1=0, weight=0, add=0)
                    elif rank_value(p):
                        emit(op=SUB_REJECT, pos=0, len=0, next0=0, next1=0, weight=0, add=0)
                    else:
                        emit(op=SUB_FREE, pos=0, len=1, next0=nxt, next1=0, weight=0, add=0)
                else:  # BRANCH
                    p = position[bit]
                    if p < n_local:
                        emit(op=SUB_BRANCH, pos=p, len=1, next0=n0, next1=n1, weight=0, add=0)
                    else:
                        emit(op=SUB_FREE, pos=0, len=1, next0=(n1 if rank_value(p) else n0), next1=0, weight=0, add=0)
        entry = heads[prog.entry]
        # qubits above the subspace's own (clean / borrowed ancillae) are post-selected on 0
        # (Subspace.project only enumerates the subspace's qubits, nodes/node.py:379)
        n_total = len(position) if n_total is None else n_total
        for q in range(prog.total_qubits, n_total):
            p = position[q]
            if p < n_local:
                entry = emit(op=SUB_ZEROS, pos=p, len=1, next0=entry, next1=0, weight=0, add=0)
            elif rank_value(p):
                entry = emit(op=SUB_REJECT, pos=0, len=0, next0=0, next1=0, weight=0, add=0)
        # peephole: jump over the no-ops (FREE with weight 0) so the device walk only executes tests
        def resolve(pc):
            while out[pc]["op"] == SUB_FREE and out[pc]["weight"] == 0:
                pc = out[pc]["next0"]
            return pc

        for r in out:
            if r["op"] in (SUB_ZEROS, SUB_FREE):
                r["next0"] = resolve(r["next0"])
            elif r["op"] == SUB_BRANCH:
                r["next0"], r["next1"] = resolve(r["next0"]), resolve(r["next1"])
        entry = resolve(entry)
        # merge chains of single-bit ZEROS on adjacent positions into one mask test
        for r in out:
            while r["op"] == SUB_ZEROS:
                nx = out[r["next0"]]
                if nx["op"] == SUB_ZEROS and nx["pos"] == r["pos"] + r["len"]:
                    r["len"] += nx["len"]
                    r["next0"] = nx["next0"]
                elif nx["op"] == SUB_ZEROS and nx["pos"] + nx["len"] == r["pos"]:
                    r["pos"], r["len"] = nx["pos"], r["len"] + nx["len"]
                    r["next0"] = nx["next0"]
                else:
                    break
        if len(out) > 65535:
            raise ValueError("subspace program too long")
        self.instrs = np.zeros(len(out), dtype=SUBINSTR_DTYPE)
        for i, r in enumerate(out):
            for k, v in r.items():
                self.instrs[k][i] = v
        self.entry = entry
        self.total_qubits = prog.total_qubits
        self.n_local = n_local
        self.rejects_all = int(self.instrs["op"][entry]) == SUB_REJECT

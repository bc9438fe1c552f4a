"""Subspace -> flat device program (``b2sv_subnode[]``), SURVEY.md App. C.

A unitaria ``Subspace`` (/root/reference/src/unitaria/subspace.py) is a tensor product of factors,
least-significant first (:117-125): ``ZeroQubitSubspace`` (one qubit fixed to 0) or
``ControlledSubspace(case_zero, case_one)`` whose top qubit picks which sub-subspace constrains the
lower qubits (:523-560).  ``test_basis`` (:297-323) walks those factors; the device does the same
walk on the flattened tree, plus the mixed-radix rank that reproduces ``enumerate_basis`` order
(:325-329) without enumerating.

Accepted inputs: a unitaria ``Subspace`` (duck-typed via ``tensor_factors`` / ``case_zero`` /
``case_one``), a string such as ``"0##"`` (most significant qubit first, subspace.py:114-125) or the
neutral nested spec used by the fixtures: a list of factors, each ``0`` or ``[spec0, spec1]``.
"""

from __future__ import annotations

import numpy as np

from .cabi import SUBNODE_DTYPE


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


class SubspaceProgram:
    """Flattened subspace: ``nodes`` (SUBNODE_DTYPE), ``root_count``, ``total_qubits``, ``dimension``."""

    def __init__(self, subspace):
        self.spec = to_spec(subspace)
        self.total_qubits = spec_total_qubits(self.spec)
        self.dimension = spec_dimension(self.spec)
        records: list = []
        self.root_count = len(self.spec)
        self._emit(self.spec, records)
        self.nodes = np.zeros(len(records), dtype=SUBNODE_DTYPE)
        for i, r in enumerate(records):
            self.nodes[i] = r

    def _emit(self, spec, records) -> int:
        """Lay out the factors of ``spec`` contiguously; children follow later.  Returns first index."""
        first = len(records)
        records.extend([None] * len(spec))
        for j, f in enumerate(spec):
            if f == 0:
                records[first + j] = (0, 1, 1, 0, (0, 0), (0, 0))
                continue
            s0, s1 = f
            if spec_total_qubits(s0) != spec_total_qubits(s1):
                raise ValueError("ControlledSubspace cases span different qubit counts")
            d0, d1 = spec_dimension(s0), spec_dimension(s1)
            c0 = self._emit(s0, records) if len(s0) else 0
            c1 = self._emit(s1, records) if len(s1) else 0
            records[first + j] = (1, 1 + spec_total_qubits(s0), d0 + d1, d0, (c0, c1), (len(s0), len(s1)))
        return first

"""Helpers shared by the CPU and GPU tests."""

from __future__ import annotations

import glob
import os

import numpy as np

from unitaria_b200.fixtures import column_error, load_circuit  # noqa: F401
from unitaria_b200.fixtures import load_columns as _load_columns

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(with_expected=None, max_qubits=None, min_qubits=None):
    names = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        if path.endswith(".cols.npz"):  # frozen columns of a case, not a case (load_columns)
            continue
        c = load_circuit(path)
        if with_expected is not None and (("expected" in c.arrays) != with_expected):
            continue
        if max_qubits is not None and max(1, c.n_qubits) > max_qubits:
            continue
        if min_qubits is not None and max(1, c.n_qubits) < min_qubits:
            continue
        names.append(os.path.splitext(os.path.basename(path))[0])
    return names


def load_golden(name):
    return load_circuit(os.path.join(GOLDEN_DIR, name + ".npz"))


def load_columns(name):
    """Frozen columns of the encoded matrix (reference ``node.compute(e_j)``, tests/golden/make_golden.py --columns)."""
    return _load_columns(os.path.join(GOLDEN_DIR, name + ".cols.npz"))


def golden_inputs(case, limit=None, rng=None):
    """Basis inputs of a golden case: [(column, basis_index)].  Vectors have the single input 0."""
    if case.meta["is_vector"]:
        return [(None, 0)]
    basis = case.arrays["basis_in"]
    cols = list(range(len(basis)))
    if limit is not None and len(cols) > limit:
        rng = rng or np.random.default_rng(0)
        cols = sorted(set([0, len(basis) - 1] + list(rng.choice(len(basis), size=limit - 2, replace=False))))
    return [(c, int(basis[c])) for c in cols]


def expected_column(case, col):
    e = case.arrays["expected"]
    return e if col is None else e[:, col]


def random_circuit(n, m, rng, kinds=None, max_controls=3):
    """Random circuit over unitaria's whole gate alphabet, built with the tequila(-like) constructors."""
    import tequila as tq

    c = tq.QCircuit()
    kinds = kinds or ["X", "Ry", "Rz", "Phase", "GlobalPhase", "H", "Zp", "SWAP", "Rx", "Y", "Xp", "Z", "Hp", "CNOT", "Toffoli"]
    for _ in range(m):
        kind = kinds[int(rng.integers(0, len(kinds)))]
        qs = [int(q) for q in rng.permutation(n)]
        nc = int(rng.integers(0, max_controls + 1))
        t, t2 = qs[0], (qs[1] if n > 1 else None)
        ctrl = qs[2 : 2 + nc]
        a = float(rng.normal())
        if kind == "X":
            c += tq.gates.X(t, control=ctrl)
        elif kind == "CNOT" and n >= 2:
            c += tq.gates.CNOT(qs[1], t)
        elif kind == "Toffoli" and n >= 3:
            c += tq.gates.Toffoli(qs[1], qs[2], t)
        elif kind == "Ry":
            c += tq.gates.Ry(a, t, control=ctrl)
        elif kind == "Rz":
            c += tq.gates.Rz(a, t, control=ctrl)
        elif kind == "Rx":
            c += tq.gates.Rx(a, t, control=ctrl)
        elif kind == "Phase":
            c += tq.gates.Phase(t, control=ctrl, angle=a)
        elif kind == "GlobalPhase":
            c += tq.gates.GlobalPhase(a).add_controls(ctrl + ([t] if rng.integers(0, 2) else []))
        elif kind == "H":
            c += tq.gates.H(t, control=ctrl)
        elif kind == "Hp":
            c += tq.gates.H(t, control=ctrl, power=a)
        elif kind == "Z":
            c += tq.gates.Z(t, control=ctrl)
        elif kind == "Zp":
            c += tq.gates.Z(t, control=ctrl, power=2.0 ** int(rng.integers(-4, 1)))
        elif kind == "Y":
            c += tq.gates.Y(t, control=ctrl)
        elif kind == "Xp":
            c += tq.gates.X(t, control=ctrl, power=a)
        elif kind == "SWAP" and n >= 2:
            c += tq.gates.SWAP(t, t2, control=ctrl)
        else:
            c += tq.gates.X(t)
    return c


def random_state(n, rng):
    psi = rng.normal(size=2**n) + 1j * rng.normal(size=2**n)
    return psi / np.linalg.norm(psi)


def dagger_gates(gates):
    """Inverse circuit of a list of GateRecords (what ``QCircuit.dagger()`` does gate by gate)."""
    from unitaria_b200.fixtures import GateRecord

    out = []
    for g in reversed(list(gates)):
        name = g.name.upper()
        if name in ("RX", "RY", "RZ", "PHASE", "GLOBALPHASE"):
            out.append(GateRecord(g.name, g.target, g.control, -g.parameter, 1.0))
        else:
            p = g.power if g.power == 1 else -g.power
            out.append(GateRecord(g.name, g.target, g.control, p, p))
    return out

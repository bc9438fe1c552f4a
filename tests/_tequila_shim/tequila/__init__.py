"""Test fixture: a minimal stand-in for the slice of the tequila API that unitaria touches.

tequila-basic (and qulacs behind it) are not installable in the build image (no network), so
this duck-typed stand-in exists to (1) let the read-only reference package import in the build
container so golden vectors can be generated from its numpy ``compute()`` path, and (2) give the
GPU tests the same gate constructors the reference uses.  It is only put on ``sys.path`` by
tests/conftest.py when the real ``tequila`` is missing.  Its ``simulate`` delegates to the CPU
oracle; it is never used by the product package.

API census it covers: SURVEY.md App. A.
"""

from __future__ import annotations

import copy as _copy
import enum
import sys
import types

import numpy as np


class BitNumbering(enum.Enum):
    MSB = 0
    LSB = 1


def _qubits(x):
    if x is None:
        return ()
    if isinstance(x, (int, np.integer)):
        return (int(x),)
    return tuple(int(i) for i in x)


class QGateImpl:
    def __init__(self, name, target, control=None):
        self.name = name
        self.target = _qubits(target)
        self.control = _qubits(control)
        if set(self.target) & set(self.control):
            raise ValueError(f"{name}: target {self.target} overlaps control {self.control}")

    @property
    def qubits(self):
        return sorted(set(self.target + self.control))

    @property
    def max_qubit(self):
        return max(self.target + self.control, default=0)

    def is_controlled(self):
        return len(self.control) > 0

    def dagger(self):
        return _copy.copy(self)

    def map_qubits(self, qubit_map):
        g = _copy.copy(self)
        g.target = tuple(qubit_map[t] for t in self.target)
        g.control = tuple(qubit_map[c] for c in self.control)
        return g

    def __repr__(self):
        s = f"{self.name}(target={self.target}"
        if self.control:
            s += f", control={self.control}"
        p = getattr(self, "parameter", None)
        if p is not None and not isinstance(self, PowerGateImpl):
            s += f", parameter={p}"
        return s + ")"


class _ParametrizedGate(QGateImpl):
    def dagger(self):
        g = _copy.copy(self)
        g.parameter = -self.parameter
        return g


class RotationGateImpl(_ParametrizedGate):
    def __init__(self, axis, angle, target, control=None):
        super().__init__("R" + axis, target, control)
        self.axis = axis
        self.parameter = angle


class PhaseGateImpl(_ParametrizedGate):
    def __init__(self, phase, target, control=None):
        super().__init__("Phase", target, control)
        self.parameter = phase


class GlobalPhaseGateImpl(_ParametrizedGate):
    def __init__(self, phase, control=None):
        super().__init__("GlobalPhase", (), control)
        self.parameter = phase


class PowerGateImpl(QGateImpl):
    def __init__(self, name, power, target, control=None):
        super().__init__(name, target, control)
        self.power = 1 if power is None else power

    @property
    def parameter(self):
        return self.power

    def dagger(self):
        g = _copy.copy(self)
        if self.power != 1:
            g.power = -self.power
        return g


class QCircuit:
    def __init__(self, gates=None):
        self.gates = list(gates) if gates is not None else []

    @property
    def qubits(self):
        s = set()
        for g in self.gates:
            s.update(g.target)
            s.update(g.control)
        return sorted(s)

    @property
    def n_qubits(self):
        return max((g.max_qubit for g in self.gates), default=0) + 1

    @property
    def depth(self):
        level = {}
        depth = 0
        for g in self.gates:
            q = g.qubits
            if not q:
                continue
            here = max(level.get(i, 0) for i in q) + 1
            for i in q:
                level[i] = here
            depth = max(depth, here)
        return depth

    def __iadd__(self, other):
        self.gates = self.gates + list(other.gates)
        return self

    def __add__(self, other):
        return QCircuit(self.gates + list(other.gates))

    def dagger(self):
        return QCircuit([g.dagger() for g in reversed(self.gates)])

    def add_controls(self, control, inpl=False):
        extra = _qubits(control)
        out = []
        for g in self.gates:
            h = _copy.copy(g)
            h.control = tuple(g.control) + tuple(x for x in extra if x not in g.control)
            if set(h.target) & set(h.control):
                raise ValueError("control overlaps target")
            out.append(h)
        return QCircuit(out)

    def map_qubits(self, qubit_map):
        return QCircuit([g.map_qubits(qubit_map) for g in self.gates])

    def __str__(self):
        return "circuit: \n" + "".join(repr(g) + "\n" for g in self.gates)

    __repr__ = __str__


def _one(g):
    return QCircuit([g])


def _power_gate(name):
    def make(target, control=None, power=None, *args, **kwargs):
        return _one(PowerGateImpl(name, power, target, control))

    make.__name__ = name
    return make


gates = types.ModuleType("tequila.gates")
gates.X = _power_gate("X")
gates.Y = _power_gate("Y")
gates.Z = _power_gate("Z")
gates.H = _power_gate("H")
gates.I = lambda target, *a, **k: _one(PowerGateImpl("I", None, target))
gates.CNOT = lambda control, target: _one(PowerGateImpl("X", None, target, control))
gates.CX = gates.CNOT
gates.Toffoli = lambda first, second, target: _one(PowerGateImpl("X", None, target, (first, second)))
gates.SWAP = lambda first, second, control=None, power=None, *a, **k: _one(
    PowerGateImpl("SWAP", power, (first, second), control)
)
gates.Rx = lambda angle, target, control=None, **k: _one(RotationGateImpl("x", angle, target, control))
gates.Ry = lambda angle, target, control=None, **k: _one(RotationGateImpl("y", angle, target, control))
gates.Rz = lambda angle, target, control=None, **k: _one(RotationGateImpl("z", angle, target, control))
gates.Phase = lambda target, control=None, angle=None, **k: _one(PhaseGateImpl(angle, target, control))
gates.GlobalPhase = lambda angle, control=None: _one(GlobalPhaseGateImpl(angle, control))
sys.modules["tequila.gates"] = gates

circuit = types.ModuleType("tequila.circuit")
circuit._gates_impl = types.ModuleType("tequila.circuit._gates_impl")
for _cls in (QGateImpl, RotationGateImpl, PhaseGateImpl, GlobalPhaseGateImpl, PowerGateImpl):
    setattr(circuit._gates_impl, _cls.__name__, _cls)
circuit.qpic = types.ModuleType("tequila.circuit.qpic")
circuit.qpic.system_has_qpic = False
circuit.QCircuit = QCircuit
sys.modules["tequila.circuit"] = circuit
sys.modules["tequila.circuit._gates_impl"] = circuit._gates_impl
sys.modules["tequila.circuit.qpic"] = circuit.qpic


class QubitWaveFunction:
    def __init__(self, array):
        self._array = array  # LSB ordered

    @staticmethod
    def from_array(array, numbering=BitNumbering.LSB):
        assert numbering == BitNumbering.LSB
        return QubitWaveFunction(np.asarray(array, dtype=np.complex128))

    @staticmethod
    def from_basis_state(n_qubits, basis_state, numbering=BitNumbering.LSB):
        assert numbering == BitNumbering.LSB
        a = np.zeros(2**n_qubits, dtype=np.complex128)
        a[basis_state] = 1
        return QubitWaveFunction(a)

    def to_array(self, numbering=BitNumbering.LSB, copy=True):
        assert numbering == BitNumbering.LSB
        return self._array.copy() if copy else self._array


# calls recorded as (n_qubits, n_gates) -- lets tests assert which seam was exercised
SIM_CALLS = []


def simulate(objective, initial_state=None, backend=None, samples=None, **kwargs):
    """Stand-in for ``tq.simulate`` (wavefunction mode only); delegates to the CPU oracle."""
    if samples is not None:
        raise NotImplementedError("sampling mode is not covered by the stand-in")
    from oracle import np_oracle  # tests put the repo root on sys.path

    n = objective.n_qubits
    if initial_state is None:
        psi0 = 0
    else:
        psi0 = initial_state.to_array(BitNumbering.LSB)
    SIM_CALLS.append((n, len(objective.gates)))
    return QubitWaveFunction(np_oracle.simulate_gates(objective.gates, n, psi0))

"""Lowering: a tequila gate list -> the flat ``b2sv_gate`` array the C ABI takes.

Input is whatever ``Circuit.simulate`` would hand to ``tq.simulate``
(/root/reference/src/unitaria/circuit.py:62-70): ``tq.QCircuit.gates``, read by duck typing
(``.name``, ``.target``, ``.control``, ``.parameter``, ``.power``), so real tequila gate objects,
the test stand-in and :class:`unitaria_b200.fixtures.GateRecord` all work.  The gate alphabet and
its conventions are the ones pinned in SURVEY.md App. A:

===========  =====================================================  ==========================
tequila      unitary on the target (applied iff all controls are 1)   lowered op
===========  =====================================================  ==========================
I            identity (padding only, circuit.py:73-78)               dropped
X            [[0,1],[1,0]]  (CNOT/Toffoli = controlled X)            ``OP_X`` (exact data move)
Y, H         Pauli-Y, Hadamard                                       ``OP_MAT1``
Z (power p)  diag(1, e^{i pi p})   (circuits/qft.py:10)              ``OP_DIAG1``
Rx, Ry       exp(-i theta/2 X|Y)  (README.md:20-24 pins Ry's sign)   ``OP_MAT1``
Rz           diag(e^{-i theta/2}, e^{+i theta/2})                    ``OP_DIAG1``
Phase        diag(1, e^{i phi})    (nodes/basic/scale.py:103)        ``OP_DIAG1``
GlobalPhase  e^{+i alpha} on the controlled block                    ``OP_PHASE``
SWAP         exchange two qubits (controlled: Fredkin)               ``OP_SWAP`` (exact data move)
===========  =====================================================  ==========================

Angles are turned into matrix entries here, in float64 on the host, so the device only multiplies.
Exact no-ops (``I``, zero angles, X.X pairs) are recognised and dropped by the planner inside the
library, which sees exact 1.0 / 0.0 entries for them.
"""

from __future__ import annotations

import math

import numpy as np

from .cabi import GATE_DTYPE, OP_DIAG1, OP_MAT1, OP_PHASE, OP_SWAP, OP_X

_SQ2 = 1.0 / math.sqrt(2.0)
_PAULI = {
    "X": ((0.0, 0.0), (1.0, 0.0), (1.0, 0.0), (0.0, 0.0)),
    "Y": ((0.0, 0.0), (0.0, -1.0), (0.0, 1.0), (0.0, 0.0)),
    "Z": ((1.0, 0.0), (0.0, 0.0), (0.0, 0.0), (-1.0, 0.0)),
    "H": ((_SQ2, 0.0), (_SQ2, 0.0), (_SQ2, 0.0), (-_SQ2, 0.0)),
}


class LoweringError(ValueError):
    pass


def _cis(angle: float) -> tuple[float, float]:
    """e^{i angle} with exact values at multiples of pi/2 (so e.g. Z^1 is exactly diag(1,-1))."""
    q = angle / (math.pi / 2)
    if q == int(q) and abs(q) < 2**40:
        return ((1.0, 0.0), (0.0, 1.0), (-1.0, 0.0), (0.0, -1.0))[int(q) % 4]
    return (math.cos(angle), math.sin(angle))


def _power_matrix(name: str, power: float):
    """U^p = P+ + e^{i pi p} P-,  P+- = (1 +- U)/2, for U with eigenvalues +-1."""
    base = _PAULI[name]
    if power == 1:
        return base
    w = complex(*_cis(math.pi * power))
    out = []
    for idx, (re, im) in enumerate(base):
        b = complex(re, im)
        eye = 1.0 if idx in (0, 3) else 0.0
        v = 0.5 * (eye + b) + w * 0.5 * (eye - b)
        out.append((v.real, v.imag))
    return tuple(out)


def _num(x, what: str) -> float:
    """Numeric gate parameter; tequila Variables / objectives (parametrised circuits) are not lowered."""
    try:
        return float(x)
    except (TypeError, ValueError) as exc:
        raise LoweringError(f"{what}: parameter {x!r} is not a number (parametrised gates are outside this engine's alphabet)") from exc


def _mask(qubits) -> int:
    m = 0
    for q in qubits:
        q = int(q)
        if q < 0 or q > 63:
            raise LoweringError(f"qubit index {q} outside [0, 63]")
        m |= 1 << q
    return m


def lower_gates(gates, n_qubits: int | None = None) -> np.ndarray:
    """Lower a sequence of tequila-like gates to a ``GATE_DTYPE`` array (one entry per non-``I`` gate).

    unitaria's circuits are built by repeating sub-circuits (QSVT alternates A and A^dagger,
    nodes/qsvt/qsvt.py:341-359; amplification repeats the whole node): the 93 675 gates of
    Pseudoinverse(BPX L=4) are 753 distinct (name, targets, controls, parameter, power) records.  Each
    distinct record is lowered once and the output is one table gather."""
    gates = gates if isinstance(gates, (list, tuple)) else list(gates)
    if len(gates) < 64:
        return _lower_each(gates, n_qubits)
    memo: dict = {}
    unique = []
    which = np.empty(len(gates), dtype=np.intp)
    for i, g in enumerate(gates):
        try:
            key = (g.name, g.target, g.control, getattr(g, "parameter", None), getattr(g, "power", None))
            j = memo.get(key)
        except TypeError:  # unhashable field (list targets, symbolic parameter): lowered on its own
            key, j = None, None
        if j is None:
            j = len(unique)
            unique.append(g)
            if key is not None:
                memo[key] = j
        which[i] = j
    table, kept = _lower_each(unique, n_qubits, keep_dropped=True)
    return table[which[kept[which]]]


def _lower_each(gates, n_qubits: int | None = None, keep_dropped: bool = False):
    """One entry per gate; ``keep_dropped``: also return the mask of entries that are real gates (not ``I``),
    with the array still holding one (zeroed) slot per input gate."""
    out = np.zeros(len(gates), dtype=GATE_DTYPE)
    kept = np.ones(len(gates), dtype=bool)
    ops, t0s, t1s, nts, ctrls, mats = out["op"], out["t0"], out["t1"], out["n_targets"], out["ctrl_mask"], out["m"]
    k = 0
    for g in gates:
        name = str(g.name)
        uname = name.upper()
        if uname == "I":
            if keep_dropped:
                kept[k] = False
                k += 1
            continue
        target = tuple(int(t) for t in (g.target or ()))
        control = tuple(int(c) for c in (g.control or ()))
        cm = _mask(control)
        if cm & _mask(target):
            raise LoweringError(f"{name}: target {target} overlaps control {control}")
        if n_qubits is not None and (cm | _mask(target)) >> max(1, n_qubits):
            raise LoweringError(f"{name}: qubit outside the {n_qubits}-qubit register")
        ctrls[k] = cm
        m = mats[k]
        if uname == "GLOBALPHASE":
            ops[k], nts[k] = OP_PHASE, 0
            m[0], m[1] = _cis(_num(g.parameter, name))
        elif uname == "SWAP":
            power = getattr(g, "power", 1)
            if power not in (1, None):
                raise LoweringError("SWAP with a power is outside unitaria's gate alphabet")
            if len(target) != 2:
                raise LoweringError("SWAP needs two targets")
            ops[k], nts[k], t0s[k], t1s[k] = OP_SWAP, 2, target[0], target[1]
        else:
            if len(target) != 1:
                raise LoweringError(f"{name}: expected one target, got {target}")
            nts[k], t0s[k] = 1, target[0]
            if uname in ("X", "Y", "H"):
                power = getattr(g, "power", 1)
                power = 1 if power is None else power
                if uname == "X" and power == 1:
                    ops[k] = OP_X
                else:
                    ops[k] = OP_MAT1
                    m[:] = [v for pair in _power_matrix(uname, _num(power, name)) for v in pair]
            elif uname == "Z":
                power = getattr(g, "power", 1)
                power = 1 if power is None else power
                ops[k] = OP_DIAG1
                m[0], m[1] = 1.0, 0.0
                m[2], m[3] = _cis(math.pi * _num(power, name))
            elif uname == "RZ":
                t = _num(g.parameter, name) / 2.0
                c, s = _cis(t)
                ops[k] = OP_DIAG1
                m[0], m[1], m[2], m[3] = c, -s, c, s
            elif uname == "PHASE":
                ops[k] = OP_DIAG1
                m[0], m[1] = 1.0, 0.0
                m[2], m[3] = _cis(_num(g.parameter, name))
            elif uname == "RY":
                t = _num(g.parameter, name) / 2.0
                c, s = _cis(t)
                ops[k] = OP_MAT1
                m[:] = [c, 0.0, -s, 0.0, s, 0.0, c, 0.0]
            elif uname == "RX":
                t = _num(g.parameter, name) / 2.0
                c, s = _cis(t)
                ops[k] = OP_MAT1
                m[:] = [c, 0.0, 0.0, -s, 0.0, -s, c, 0.0]
            else:
                raise LoweringError(f"gate {name!r} is outside unitaria's gate alphabet (SURVEY.md App. A)")
        k += 1
    if keep_dropped:
        return out, kept
    return out[:k]


def touches_any_qubit(gates) -> bool:
    """``len(tq_circuit.qubits) != 0`` (circuit.py:54) without needing a QCircuit object."""
    for g in gates:
        if (g.target and len(g.target)) or (g.control and len(g.control)):
            return True
    return False

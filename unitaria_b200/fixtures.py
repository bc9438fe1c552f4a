"""Gate-list (de)serialisation: tequila-level circuits as compact ``.npz`` files.

A saved circuit keeps the *tequila-level* description of every gate (name, targets, controls,
angle/power) -- i.e. what ``tq.QCircuit.gates`` holds when ``Circuit.simulate`` is called
(/root/reference/src/unitaria/circuit.py:62-70) -- so a circuit built once by unitaria's Python
(1-20 s for the large block encodings) can be replayed without rebuilding
it.  ``GateRecord`` duck-types a tequila gate, so :mod:`unitaria_b200.lowering` consumes loaded
circuits exactly like live tequila objects.
"""

from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

GATE_NAMES = ("I", "X", "Y", "Z", "H", "Rx", "Ry", "Rz", "Phase", "GlobalPhase", "SWAP")
_NAME_INDEX = {n.upper(): i for i, n in enumerate(GATE_NAMES)}
_POWER_GATES = {"I", "X", "Y", "Z", "H", "SWAP"}


@dataclass
class GateRecord:
    """Duck-types the attributes of a tequila gate that unitaria and the lowering read."""

    name: str
    target: tuple
    control: tuple
    parameter: float  # angle; for power gates the power
    power: float = 1.0

    def __repr__(self):
        return f"{self.name}(target={self.target}, control={self.control}, parameter={self.parameter})"


@dataclass
class SavedCircuit:
    gates: list
    n_qubits: int
    meta: dict = field(default_factory=dict)
    arrays: dict = field(default_factory=dict)


def encode_gates(gates) -> dict:
    m = len(gates)
    kind = np.zeros(m, dtype=np.uint8)
    t0 = np.full(m, -1, dtype=np.int16)
    t1 = np.full(m, -1, dtype=np.int16)
    ctrl = np.zeros(m, dtype=np.uint64)
    param = np.zeros(m, dtype=np.float64)
    for j, g in enumerate(gates):
        uname = g.name.upper()
        kind[j] = _NAME_INDEX[uname]
        tgt = tuple(g.target or ())
        if len(tgt) > 0:
            t0[j] = tgt[0]
        if len(tgt) > 1:
            t1[j] = tgt[1]
        mask = 0
        for c in g.control or ():
            mask |= 1 << int(c)
        ctrl[j] = mask
        if uname in _POWER_GATES:
            p = getattr(g, "power", 1)
            param[j] = 1.0 if p is None else float(p)
        else:
            param[j] = float(g.parameter)
    return {"gate_kind": kind, "gate_t0": t0, "gate_t1": t1, "gate_ctrl": ctrl, "gate_param": param}


def decode_gates(arrays) -> list:
    kind, t0, t1 = arrays["gate_kind"], arrays["gate_t0"], arrays["gate_t1"]
    ctrl, param = arrays["gate_ctrl"], arrays["gate_param"]
    out = []
    for j in range(len(kind)):
        name = GATE_NAMES[int(kind[j])]
        target = tuple(int(t) for t in (t0[j], t1[j]) if t >= 0)
        mask = int(ctrl[j])
        control = tuple(q for q in range(64) if (mask >> q) & 1)
        p = float(param[j])
        out.append(GateRecord(name, target, control, p, p if name.upper() in _POWER_GATES else 1.0))
    return out


def save_circuit(path, gates, n_qubits: int, meta: dict | None = None, **arrays) -> None:
    payload = encode_gates(gates)
    payload["n_qubits"] = np.int64(n_qubits)
    payload["meta_json"] = np.array(json.dumps(meta or {}))
    for k, v in arrays.items():
        payload["x_" + k] = np.asarray(v)
    np.savez_compressed(path, **payload)


def load_circuit(path) -> SavedCircuit:
    with np.load(path, allow_pickle=False) as z:
        gates = decode_gates(z)
        meta = json.loads(str(z["meta_json"]))
        arrays = {k[2:]: z[k] for k in z.files if k.startswith("x_")}
        return SavedCircuit(gates, int(z["n_qubits"]), meta, arrays)


def load_columns(path):
    """Frozen columns of an encoded matrix (the reference's ``node.compute(e_j)``, written by
    tests/golden/make_golden.py --columns): ``[(basis index of input j, nonzero indices, nonzero values)]`` and the
    output dimension; ``(None, 0)`` if the file does not exist."""
    import os

    if not os.path.exists(path):
        return None, 0
    with np.load(path) as z:
        ptr, idx, val = z["col_ptr"], z["col_index"], z["col_value"]
        cols = [(int(b), idx[ptr[k] : ptr[k + 1]], val[ptr[k] : ptr[k + 1]]) for k, b in enumerate(z["col_basis"])]
        return cols, int(z["dim_out"])


def column_error(got, indices, values) -> float:
    """max |got - column| of a dense result against a sparse expected column."""
    err_nz = float(np.abs(got[indices] - values).max()) if len(indices) else 0.0
    mask = np.ones(len(got), dtype=bool)
    mask[indices] = False
    rest = got[mask]
    return max(err_nz, float(np.abs(rest).max()) if rest.size else 0.0)

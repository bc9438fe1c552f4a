"""Cost of the in-tile op kinds: one tile sweep of G gates on a 26-qubit complex128 register."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "_tequila_shim"))
import tequila as tq
from unitaria_b200 import PlanOpts, StateVector, lower_gates

n = 26
rng = np.random.default_rng(0)

def circ(kind, m):
    c = tq.QCircuit()
    for i in range(m):
        a = float(rng.normal())
        if kind == "ry_same":            # m uncontrolled Ry on one qubit, separated by CNOTs from another (mux-able)
            c += tq.gates.Ry(a, 4) + tq.gates.CNOT(9, 4)
        elif kind == "ry_q3":            # full-touch dense gate, low target
            c += tq.gates.Ry(a, 3) + tq.gates.Rz(a, 5)   # alternate targets: no fusion by mux
        elif kind == "ry_q10":
            c += tq.gates.Ry(a, 10) + tq.gates.Rz(a, 11)
        elif kind == "blk2":             # gates confined to {3,4}
            c += tq.gates.Ry(a, 3) + tq.gates.CNOT(3, 4) + tq.gates.Rz(a, 4)
        elif kind == "blk3":
            c += tq.gates.Ry(a, 3) + tq.gates.CNOT(3, 4) + tq.gates.Rz(a, 4) + tq.gates.CNOT(4, 5) + tq.gates.Ry(a, 5)
        elif kind == "cx":               # permutation ops, one control
            c += tq.gates.CNOT(3, 8) + tq.gates.CNOT(8, 2)
        elif kind == "ctrl3":            # heavily controlled dense gates (local controls)
            c += tq.gates.Ry(a, 4, control=[0, 1, 2]) + tq.gates.Ry(a, 9, control=[5, 6, 7])
    return c

with StateVector(n) as sv:
    for kind, m in [("ry_q3", 8), ("ry_q10", 8), ("cx", 8), ("ctrl3", 8), ("ry_same", 16), ("blk2", 8), ("blk3", 6)]:
        c = circ(kind, m)
        low = lower_gates(c.gates, n)
        for label, kw in [("plain", dict(mux=False, perm_runs=False)), ("fused", dict(perm_runs=False))]:
            opts = PlanOpts.make(timing=True, **kw)
            for _ in range(2):
                sv.set_basis(1); sv.apply_lowered(low, opts)
            sv.stats_reset()
            for _ in range(5):
                sv.apply_lowered(low, opts)
            st = sv.stats()
            L = st["launches_by_class"]["tile"]
            print(f"{kind:8s} {label:6s} gates {len(low):3d}  tile launches/apply {L/5:4.1f}  ms/apply {st['ms_by_class']['tile']/5:7.3f}  (HBM floor {2*16*2**n/6551e9*1e3*L/5:.3f})", flush=True)

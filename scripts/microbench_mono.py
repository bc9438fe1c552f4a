"""Cost of monomial phases: one tile sweep of G X-family gates on a 26-qubit complex128 register,
with and without monomial phases (and without any phases)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "_tequila_shim"))
import tequila as tq
from unitaria_b200 import PlanOpts, StateVector, lower_gates

n = int(os.environ.get("MB_QUBITS", "26"))
dtype = os.environ.get("MB_DTYPE", "complex128")
rng = np.random.default_rng(0)

def circ(kind, m, nq):
    """m gates whose targets and controls all lie in qubits [0, nq)"""
    c = tq.QCircuit()
    for i in range(m):
        qs = [int(q) for q in rng.permutation(nq)]
        if kind == "x":
            c += tq.gates.X(qs[0])
        elif kind == "cx":
            c += tq.gates.CNOT(qs[1], qs[0])
        elif kind == "ccx":
            c += tq.gates.Toffoli(qs[1], qs[2], qs[0])
        elif kind == "c3x":
            c += tq.gates.X(qs[0], control=qs[1:4])
        elif kind == "mixed":   # what ripple-carry arithmetic looks like
            k = int(rng.integers(0, 4))
            c += tq.gates.X(qs[0], control=qs[1:1 + k])
        elif kind == "phase":
            c += tq.gates.Phase(qs[0], control=qs[1:2], angle=float(rng.normal())) + tq.gates.CNOT(qs[2], qs[3])
    return c

with StateVector(n, dtype) as sv:
    floor = 2 * sv.amp_bytes * 2**n / 6551e9 * 1e3
    for kind, m, nq in [("x", 16, 8), ("x", 64, 8), ("cx", 16, 8), ("cx", 64, 8), ("cx", 64, 10), ("cx", 64, 11), ("ccx", 64, 8), ("ccx", 64, 10),
                        ("c3x", 64, 10), ("mixed", 74, 10), ("mixed", 300, 10), ("phase", 32, 8), ("phase", 32, 10)]:
        c = circ(kind, m, nq)
        low = lower_gates(c.gates, n)
        for label, kw in [("mono", dict(perm_runs=False)), ("no-mono", dict(perm_runs=False, mono=False)), ("no-phases", dict(perm_runs=False, phases=False))]:
            opts = PlanOpts.make(timing=True, **kw)
            for _ in range(2):
                sv.set_basis(1); sv.apply_lowered(low, opts)
            sv.stats_reset()
            for _ in range(4):
                sv.apply_lowered(low, opts)
            st = sv.stats()
            L = st["launches_by_class"]["tile"]
            ms = st["ms_by_class"]["tile"] / 4
            print(f"{kind:6s} m={len(low):3d} nq={nq:2d} {label:9s} tile launches/apply {L/4:4.1f}  ms/apply {ms:7.3f}  = {ms/floor:5.2f} sweeps ({(ms - floor*L/4)/floor/len(low):.4f} sweeps/gate over the floor)", flush=True)

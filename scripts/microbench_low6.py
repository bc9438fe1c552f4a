"""Cost model of the low-qubit warp-shuffle sweep: G gates on qubits 0..5 of a 28-qubit complex128 register."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "_tequila_shim"))
import tequila as tq
from unitaria_b200 import PlanOpts, StateVector, lower_gates

n = 28
rng = np.random.default_rng(0)
floor = 2 * 16 * 2**n / 6551e9 * 1e3
with StateVector(n) as sv:
    for kind in ("ry_t5", "ry_t2", "x_t2", "rz_t2", "mux_t5", "mux_t0"):
        for m in (1, 4, 16):
            c = tq.QCircuit()
            for i in range(m):
                a = float(rng.normal())
                if kind == "ry_t5":
                    c += tq.gates.Ry(a, 5) + tq.gates.Rz(a, 4)
                elif kind == "ry_t2":
                    c += tq.gates.Ry(a, 2) + tq.gates.Rz(a, 4)
                elif kind == "x_t2":
                    c += tq.gates.CNOT(3, 2) + tq.gates.Rz(a, 4)
                elif kind == "rz_t2":
                    c += tq.gates.Rz(a, 2) + tq.gates.Phase(4, angle=a)
                elif kind == "mux_t5":
                    for j in range(4):
                        c += tq.gates.Ry(a + j, 5) + tq.gates.CNOT(j, 5)
                    c += tq.gates.Rz(a, 4)
                elif kind == "mux_t0":
                    for j in range(1, 5):
                        c += tq.gates.Ry(a + j, 0) + tq.gates.CNOT(j, 0)
                    c += tq.gates.Rz(a, 4)
            low = lower_gates(c.gates, n)
            opts = PlanOpts.make(timing=True)
            for _ in range(2):
                sv.set_basis(1); sv.apply_lowered(low, opts)
            sv.stats_reset()
            for _ in range(5):
                sv.apply_lowered(low, opts)
            st = sv.stats()
            L = st["launches_by_class"]
            print(f"{kind:7s} m={m:2d} gates {len(low):3d} launches {dict((k, v // 5) for k, v in L.items() if v)} low6 ms/apply {st['ms_by_class']['low6']/5:7.3f} tile {st['ms_by_class']['tile']/5:7.3f} (HBM floor {floor:.3f})", flush=True)

"""Marginal cost of one shared-memory pass of each kind inside a compute-bound tile sweep.

Base: a chain of NB dense 8x8 blocks on overlapping qubit triples of the tile ({0,1,2}, {2,3,4}, ...: neighbours share one
qubit, so nothing merges and nothing commutes) -- one sweep, NB passes.  Variants put one extra pass of a kind between every
two blocks; (variant - base) / (NB - 1) is what such a pass costs once the sweep is bound by its passes.
Usage: python scripts/microbench_passes.py [--qubits 28] [--describe]   (--describe: plans only, no GPU)
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unitaria_b200 import PlanOpts, cabi, lower_gates  # noqa: E402
from unitaria_b200.fixtures import GateRecord  # noqa: E402

KIND = {0: "MAT1", 1: "DIAG1", 2: "X", 3: "SWAP", 4: "PHASE", 5: "MUX", 6: "BLOCK", 7: "REGPHASE", 8: "MONO", 9: "NOP"}
ap = argparse.ArgumentParser()
ap.add_argument("--qubits", type=int, default=28)
ap.add_argument("--blocks", type=int, default=10)
ap.add_argument("--describe", action="store_true")
args = ap.parse_args()
n, NB = args.qubits, args.blocks
rng = np.random.default_rng(0)


def G(name, t, c=(), p=None):
    return GateRecord(name, tuple(t), tuple(c), p, 1.0)


def block(j):
    a, b, c = (2 * j) % 10, (2 * j + 1) % 10, (2 * j + 2) % 10
    r = [float(x) for x in rng.normal(size=6)]
    return [G("Ry", [a], (), r[0]), G("Rz", [b], (), r[1]), G("X", [b], [a]), G("Rx", [c], (), r[2]), G("X", [c], [b]), G("Rz", [c], (), r[3]),
            G("Ry", [b], (), r[4]), G("X", [a], [c]), G("Rx", [a], (), r[5])]


def extra(kind, j):
    x, y = 10, 11  # tile qubits the blocks never touch
    o = (2 * j + 5) % 10  # a block qubit of neither neighbour
    if kind == "x2":
        return [G("X", [x], [y, o])]
    if kind == "x0":
        return [G("X", [x])]
    if kind == "mat1":
        return [G("Ry", [x], (), 0.3)]
    if kind == "swap1":
        return [G("SWAP", [x, y], [o])]
    if kind == "regphase2":
        return [G("Ry", [x], (), 0.3), G("Rz", [y], (), 0.7)]
    if kind == "regphase4c":
        return [G("Ry", [x], [o], 0.3), G("Rz", [y], [o], 0.7), G("X", [x], [y]), G("Ry", [y], [o], 0.2)]
    if kind == "mono8":
        return [G("X", [x], [y]), G("X", [y], [o]), G("X", [o], [x]), G("SWAP", [x, y]), G("X", [x], [o]), G("X", [y], [x]), G("X", [o], [y]), G("X", [x])]
    return []


def circuit(kind):
    gates = []
    for j in range(NB):
        gates += block(j)
        if j + 1 < NB:
            gates += extra(kind, j)
    return gates


kinds = ["base", "x2", "x0", "mat1", "swap1", "regphase2", "regphase4c", "mono8"]
emul = None
path = os.path.join(ROOT, "tests", "emul", "libtile_emul.so")
if os.path.exists(path):
    emul = ctypes.CDLL(path)
rows = {}
sv = None
if not args.describe:
    from unitaria_b200 import StateVector

    sv = StateVector(n)
for kind in kinds:
    rng = np.random.default_rng(0)
    low = np.ascontiguousarray(lower_gates(circuit(kind), n), dtype=cabi.GATE_DTYPE)
    opts = PlanOpts.make(timing=True, jit="off")
    segs, _ = cabi.plan_describe(n, cabi.B2SV_C128, low, opts)[:2]
    desc = []
    if emul is not None:
        out = np.zeros(4096, dtype=np.int32)
        for si, s in enumerate(segs):
            if s[0] != cabi.SEG_TILE:
                desc.append(f"[cls {s[0]}]")
                continue
            m = emul.emul_describe_segment(n, cabi.B2SV_C128, low.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(low)), ctypes.byref(opts), si, out.ctypes.data_as(ctypes.c_void_p), len(out))
            desc.append(" ".join(f"{KIND.get(int(out[2*i]), out[2*i])}:{int(out[2*i+1])}" for i in range(max(m, 0))))
    ms = None
    if sv is not None:
        for _ in range(2):
            sv.fill_random(seed=1, support_bits=n)
            sv.apply_lowered(low, opts)
        sv.stats_reset()
        reps = 5
        for _ in range(reps):
            sv.apply_lowered(low, opts)
        st = sv.stats()
        ms = sum(st["ms_by_class"].values()) / reps
    rows[kind] = (ms, len(segs), desc)
    print(kind, "ms", None if ms is None else round(ms, 4), "segments", len(segs), "|", " || ".join(desc)[:400], flush=True)
if sv is not None:
    base = rows["base"][0]
    hbm = 2 * 16 * 2.0**n / 6551.4e9 * 1e3
    print(f"HBM floor of one sweep at {n} qubits: {hbm:.3f} ms; base = {NB} block passes: {base:.3f} ms")
    for kind in kinds[1:]:
        extra_passes = max(1, sum(len(d.split()) for d in rows[kind][2]) - NB)
        per = (rows[kind][0] - base) / extra_passes
        print(f"  {kind:11s}: {extra_passes:2d} extra passes, {per:7.4f} ms each  (= {per / hbm:5.2f} HBM sweeps; x4 at 30 qubits)")
    sv.close()

#!/usr/bin/env python
"""Run `reps` steps (set_basis + circuit + post-selected norm) of one frozen workload: the short command the
ncu captures wrap (scripts/gpu_r2_*.sh).  Prints the tile-launch index of every tile segment so that a
capture window (-s / -c with -k regex:k_tile_swz) can be aimed at chosen sweeps."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unitaria_b200 import PlanOpts, StateVector, SubspaceProgram, cabi, lower_gates  # noqa: E402
from unitaria_b200.fixtures import load_circuit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--dtype", default="c128")
ap.add_argument("--qubits", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--jit", default="off")
ap.add_argument("--dense", action="store_true", help="seeded dense input instead of a basis state")
ap.add_argument("--list", action="store_true", help="only print the plan (no GPU needed)")
args = ap.parse_args()
case = load_circuit(os.path.join(ROOT, "tests", "golden", args.workload + ".npz"))
n = max(1, case.n_qubits, args.qubits)
low = lower_gates(case.gates, n)
dt = cabi.B2SV_C128 if args.dtype == "c128" else cabi.B2SV_C64
opts = PlanOpts.make(timing=True, jit=args.jit)
segs, _, common = cabi.plan_describe(n, dt, low, opts, with_common=True)
k = 0
for si, (cls, mask, idx) in enumerate(segs):
    if cls == cabi.SEG_TILE:
        print(f"seg {si} tile_launch {k} gates {len(idx)} common {bin(common[si][0]).count('1')}")
        k += 1
if args.list:
    sys.exit(0)
prog = SubspaceProgram(case.meta["subspace_out"])
with StateVector(n, "complex128" if args.dtype == "c128" else "complex64") as sv:
    for r in range(args.reps):
        if args.dense:
            import numpy as np

            rng = np.random.default_rng(0)
            sv.upload_random(0) if hasattr(sv, "upload_random") else sv.upload((rng.normal(size=2**n) + 1j * rng.normal(size=2**n)) / 2 ** (n / 2 + 0.5))
        else:
            sv.set_basis(r)
        sv.apply_lowered(low, opts)
        print("norm2", sv.project_norm2(prog))
    st = sv.stats()
    print({c: (st["launches_by_class"][c], round(st["ms_by_class"][c], 3)) for c in st["ms_by_class"] if st["launches_by_class"][c]})

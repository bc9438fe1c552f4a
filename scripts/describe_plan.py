#!/usr/bin/env python
"""Print the device op program of chosen tile segments of a frozen workload (host only, uses the test emulator
library tests/emul/libtile_emul.so): kind/detail per op pass, as the tile kernels would execute it."""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unitaria_b200 import PlanOpts, cabi, lower_gates  # noqa: E402
from unitaria_b200.fixtures import load_circuit  # noqa: E402

KIND = {0: "MAT1", 1: "DIAG1", 2: "X", 3: "SWAP", 4: "PHASE", 5: "MUX", 6: "BLOCK", 7: "REGPHASE", 8: "MONO", 9: "NOP"}
ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--dtype", default="c128")
ap.add_argument("--qubits", type=int, default=0)
ap.add_argument("--segs", default="", help="comma-separated segment indices (default: all tile segments)")
args = ap.parse_args()
case = load_circuit(os.path.join(ROOT, "tests", "golden", args.workload + ".npz"))
n = max(1, case.n_qubits, args.qubits)
low = np.ascontiguousarray(lower_gates(case.gates, n), dtype=cabi.GATE_DTYPE)
dt = cabi.B2SV_C128 if args.dtype == "c128" else cabi.B2SV_C64
opts = PlanOpts.make(jit="off")
lib = ctypes.CDLL(os.path.join(ROOT, "tests", "emul", "libtile_emul.so"))
segs, _, common = cabi.plan_describe(n, dt, low, opts, with_common=True)
want = [int(s) for s in args.segs.split(",") if s] or [i for i, s in enumerate(segs) if s[0] == cabi.SEG_TILE]
out = np.zeros(4096, dtype=np.int32)
for si in want:
    m = lib.emul_describe_segment(n, dt, low.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(low)), ctypes.byref(opts), si, out.ctypes.data_as(ctypes.c_void_p), len(out))
    tile = [q for q in range(64) if (segs[si][1] >> q) & 1]
    com = [q for q in range(64) if (common[si][0] >> q) & 1]
    print(f"seg {si}: gates {len(segs[si][2])} tile {tile} common {com} passes {m}: " + " ".join(f"{KIND.get(int(out[2*i]), out[2*i])}:{int(out[2*i+1])}" for i in range(max(m, 0))))

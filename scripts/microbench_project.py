#!/usr/bin/env python
"""Time the post-selection norm (k_project_norm2) for a one-ancilla subspace with the ancilla at different index positions,
through the single-register program (b2sv_project_norm2) and through a per-rank specialised one (b2sv_project_norm2_at)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unitaria_b200 import StateVector, SubspaceProgram  # noqa: E402
from unitaria_b200.subspace_prog import ShardProgram  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
with StateVector(n, "complex128") as sv:
    sv.fill_random(1) if hasattr(sv, "fill_random") else sv.set_basis(0)
    for p in (0, 1, 2, 3, 4, 5, 8, 12, 20, n - 2, n - 1):
        spec = "".join("0" if q == p else "#" for q in range(n))[::-1]  # string is MSB first
        prog = SubspaceProgram(spec)
        shard = ShardProgram(prog, list(range(n)), n, lambda pos: 0, n)
        for label, fn in (("single", lambda: sv.project_norm2(prog)), ("shard", lambda: sv.project_norm2_shard(shard))):
            fn()
            t0 = time.perf_counter()
            for _ in range(5):
                v = fn()
            dt = (time.perf_counter() - t0) / 5
            gb = 16 * 2 ** (n - 1) / 1e9
            print(f"ancilla at {p:2d} {label:6s}: {dt*1e3:7.3f} ms  {gb/dt/1e3:6.2f} TB/s of member bytes  norm2 {v:.6f}")

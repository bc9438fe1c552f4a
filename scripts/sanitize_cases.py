#!/usr/bin/env python
"""Small-register runs of every kernel family, for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py

Each case is checked against the CPU oracle as well, so a sanitizer-clean run is also a correct one.  Registers stay
at <= 15 qubits: the tools slow kernels down by one to two orders of magnitude."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "_tequila_shim"))
from oracle import c_oracle, np_oracle  # noqa: E402  (checker only)
from tests.helpers import load_golden, random_circuit, random_state  # noqa: E402
from unitaria_b200 import PlanOpts, StateVector, SubspaceProgram, lower_gates  # noqa: E402

rng = np.random.default_rng(0)
done = []


def check(name, got, want, tol=1e-12):
    err = float(np.abs(got - want).max())
    assert err <= tol, (name, err)
    done.append(f"{name}: max err {err:.1e}")


# tile sweeps with register phases, monomial phases (walk tables), dense blocks in the constant bank, multiplexors
for name in ("bpx_L3", "gaussian_s3_t8", "laplace_n12", "pinv_bpx_L2"):
    case = load_golden(name)
    n = case.n_qubits
    gates = case.gates[:6000]
    psi0 = random_state(n, rng)
    want = c_oracle.simulate_gates(gates, n, psi0)
    for label, kw in (("default", {}), ("tiles_only", dict(perm_runs=False)), ("small_tiles", dict(tile_qubits=5, lookahead=4)), ("jit_sync", dict(jit="sync"))):
        if len(gates) > 3000 and label != "default":
            continue
        for dtype, tol in (("complex128", 1e-12), ("complex64", 2e-4)):
            with StateVector(n, dtype) as sv:
                sv.upload(psi0)
                sv.apply_lowered(lower_gates(gates, n), PlanOpts.make(**kw))
                check(f"{name}/{label}/{dtype}", sv.download(), want, tol)
# permutation-heavy random circuits: monomial phases with factors, batched op programs
for n, seed in ((9, 1), (13, 2), (15, 3)):
    r = np.random.default_rng(seed)
    kinds = ["X"] * 8 + ["CNOT"] * 3 + ["Toffoli"] * 3 + ["SWAP"] * 3 + ["Phase", "GlobalPhase", "Z", "Zp", "Rz"] * 2 + ["Ry", "H"]
    c = random_circuit(n, 500, r, kinds=kinds, max_controls=4)
    psi0 = random_state(n, r)
    want = c_oracle.simulate_gates(c.gates, n, psi0)
    for kw in (dict(perm_runs=False), dict(), dict(fuse=False)):
        with StateVector(n) as sv:
            sv.upload(psi0)
            sv.apply(c.gates, PlanOpts.make(**kw))
            check(f"random{n}/{kw}", sv.download(), want)
# lazy initialisation, live window + scatter form of the permutation sweep, low-qubit warp-shuffle sweep, projections
for name in ("laplace_n12", "cfg5_bits5", "gaussian_s3_t8"):
    case = load_golden(name)
    n = case.n_qubits
    prog = SubspaceProgram(case.meta["subspace_out"])
    members = np_oracle.enumerate_basis(case.meta["subspace_out"])
    with StateVector(n) as sv:
        for b in (0, 3, (1 << n) - 1):
            want = c_oracle.simulate_gates(case.gates, n, b)
            sv.set_basis(b)
            sv.apply(case.gates, PlanOpts.make(jit="sync"))
            n2 = sv.project_norm2(prog)
            check(f"{name}/basis{b}/norm2", np.array([n2]), np.array([np.sum(np.abs(want[members]) ** 2)]))
            check(f"{name}/basis{b}/gather", sv.project(prog, 2.0), 2.0 * want[members])
            check(f"{name}/basis{b}/state", sv.download(), want)
        assert np.array_equal(sv.enumerate_basis(prog), members)
        vec = rng.normal(size=len(members)) + 0j
        sv.embed(prog, vec)
        check(f"{name}/embed", sv.project(prog), vec, 0)
print("\n".join(done))
print(f"sanitize_cases: {len(done)} checks passed")

"""Fusion planner (host C++ inside the library, exercised through b2sv_plan_describe; CPU only).

A plan is valid iff executing its segments in order -- each segment's gates in the listed order --
gives the same state as the original gate order.  The CPU oracle executes both orders.
"""

import numpy as np
import pytest

from oracle import np_oracle
from tests.helpers import golden_names, load_golden, random_circuit, random_state
from unitaria_b200 import cabi
from unitaria_b200.lowering import lower_gates


def popcount(x):
    return bin(x).count("1")


RELABEL: dict = {}


def plan_for(gates, n, dtype=cabi.B2SV_C128, **kw):
    live = [g for g in gates if g.name != "I"]
    low = lower_gates(live, n)
    segs, scalar, _common, relabel = cabi.plan_describe(max(1, n), dtype, low, cabi.PlanOpts.make(**kw), with_relabel=True)
    RELABEL[id(segs)] = relabel
    return live, low, segs, scalar


def replay(live, segs, scalar, n, psi0):
    from unitaria_b200.fixtures import GateRecord

    relabel = RELABEL.get(id(segs), {})

    def gate(i):
        # gates the SWAP frame runs on renamed qubits (and its synthetic layout-restoring SWAPs, ids >= len(live))
        if i not in relabel:
            return live[i]
        op, t0, t1, ctrl = relabel[i]
        controls = tuple(q for q in range(64) if (ctrl >> q) & 1)
        if i >= len(live):
            assert op == cabi.OP_SWAP
            return GateRecord("SWAP", (t0, t1), controls, 1.0, 1.0)
        g = live[i]
        target = (t0, t1) if g.name.upper() == "SWAP" else ((t0,) if len(g.target) else ())
        return GateRecord(g.name, target, controls, g.parameter, getattr(g, "power", 1.0))

    order = [gi for _cls, _mask, idx in segs for gi in idx]
    # global phases without controls are folded into `scalar`; dropped no-ops, X.X pairs and renamed SWAPs vanish
    psi = np_oracle.simulate_gates([gate(i) for i in order], n, psi0)
    return scalar * psi


@pytest.mark.parametrize("n,seed", [(3, 0), (6, 1), (9, 2), (12, 3), (14, 4)])
@pytest.mark.parametrize("lookahead", [1, 96])
def test_plan_is_a_valid_reordering_random(n, seed, lookahead):
    rng = np.random.default_rng(seed)
    circuit = random_circuit(n, 400, rng)
    live, low, segs, scalar = plan_for(circuit.gates, n, lookahead=lookahead, tile_qubits=5 if n > 6 else 0)
    psi0 = random_state(n, rng)
    want = np_oracle.simulate_gates(live, n, psi0)
    got = replay(live, segs, scalar, n, psi0)
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("name", ["laplace_n8", "gaussian_s3_t8", "bpx_L3", "cfg5_bits5", "laplace_readme_n8", "laplace_n12"])
def test_plan_is_a_valid_reordering_goldens(name):
    case = load_golden(name)
    n = case.n_qubits
    rng = np.random.default_rng(5)
    live, low, segs, scalar = plan_for(case.gates, n, tile_qubits=6)
    psi0 = random_state(n, rng)
    want = np_oracle.simulate_gates(live, n, psi0)
    got = replay(live, segs, scalar, n, psi0)
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("name", golden_names(min_qubits=13))
def test_plan_structure(name):
    case = load_golden(name)
    n = case.n_qubits
    live, low, segs, _ = plan_for(case.gates, n)
    k = 12
    seen = set()
    relabel = RELABEL[id(segs)]

    def planned(gi):
        """(op, t0, t1, m) of the gate the plan executes under source id gi (the SWAP frame may have renamed its qubits)."""
        if gi in relabel:
            op, t0, t1, _ctrl = relabel[gi]
            return op, t0, t1, (low[gi]["m"] if gi < len(low) else np.zeros(8))
        rec = low[gi]
        return int(rec["op"]), int(rec["t0"]), int(rec["t1"]), rec["m"]

    for cls, mask, idx in segs:
        assert cls in (cabi.SEG_GATE1Q, cabi.SEG_TILE, cabi.SEG_PERM, cabi.SEG_LOW6)
        assert idx and not (set(idx) & seen)
        seen.update(idx)
        if cls == cabi.SEG_LOW6:
            assert mask == 0x3F
        if cls == cabi.SEG_TILE:
            assert popcount(mask) == min(k, n) and mask < (1 << n)
            for gi in idx:
                op, t0, t1, m = planned(gi)
                if op in (cabi.OP_MAT1, cabi.OP_X):
                    # a DIAG1-able MAT1 may sit outside the tile; a genuinely dense one may not
                    dense = op == cabi.OP_X or not (m[2] == 0 and m[3] == 0 and m[4] == 0 and m[5] == 0)
                    if dense:
                        assert (mask >> t0) & 1, (name, gi)
                elif op == cabi.OP_SWAP:
                    assert (mask >> t0) & 1 and (mask >> t1) & 1
        if cls == cabi.SEG_PERM:
            assert all(planned(gi)[0] in (cabi.OP_X, cabi.OP_SWAP) for gi in idx)
    assert len([gi for gi in seen if gi < len(live)]) <= len(live)


def test_headline_circuit_needs_few_sweeps():
    """cfg 2 (Laplace n=26, 29 qubits): 1450 tequila gates must collapse to a handful of sweeps."""
    case = load_golden("cfg2_laplace_n26")
    live, low, segs, _ = plan_for(case.gates, case.n_qubits)
    n_perm = sum(1 for s in segs if s[0] == cabi.SEG_PERM)
    assert n_perm >= 1
    assert len(segs) <= 24, len(segs)
    live2, low2, segs2, _ = plan_for(case.gates, case.n_qubits, perm_runs=False)
    assert len(segs2) > 4 * len(segs)


def test_unfused_plan_is_one_launch_per_live_gate():
    case = load_golden("laplace_n3")
    live, low, segs, scalar = plan_for(case.gates, case.n_qubits, fuse=False)
    assert all(cls == cabi.SEG_GATE1Q and len(idx) == 1 for cls, _m, idx in segs)
    order = [idx[0] for _c, _m, idx in segs]
    assert order == sorted(order)


def test_noops_and_xx_pairs_are_dropped():
    import tequila as tq

    c = tq.gates.X(2) + tq.gates.X(2) + tq.gates.Ry(0.0, 1) + tq.gates.CNOT(0, 1) + tq.gates.CNOT(0, 1) + tq.gates.GlobalPhase(0.5)
    c += tq.gates.Ry(0.3, 0)
    live, low, segs, scalar = plan_for(c.gates, 3)
    planned = [gi for _c, _m, idx in segs for gi in idx]
    assert planned == [len(live) - 1]
    assert abs(scalar - np.exp(0.5j)) < 1e-15


def test_permutation_run_specialisation_compiles_with_nvrtc():
    """The run-time specialised kernel source for config 2's 1434-gate run compiles for sm_100a
    (NVRTC needs no GPU)."""
    import ctypes

    case = load_golden("cfg2_laplace_n26")
    low = lower_gates(case.gates, case.n_qubits)
    perm = low[(low["op"] == cabi.OP_X) | (low["op"] == cabi.OP_SWAP)].copy()
    size = ctypes.c_size_t()
    lib = cabi.load()
    for dtype in (cabi.B2SV_C128, cabi.B2SV_C64):
        rc = lib.b2sv_perm_jit_compile_check(case.n_qubits, dtype, perm.ctypes.data, len(perm), ctypes.byref(size))
        assert rc == 0, lib.b2sv_last_error()
        assert size.value > 10000
    # wide registers (sharded runs): 34 qubits
    case = load_golden("cfg5_bits27")
    low = lower_gates(case.gates, case.n_qubits)
    perm = low[(low["op"] == cabi.OP_X) | (low["op"] == cabi.OP_SWAP)].copy()
    rc = lib.b2sv_perm_jit_compile_check(case.n_qubits, cabi.B2SV_C128, perm.ctypes.data, len(perm), ctypes.byref(size))
    assert rc == 0, lib.b2sv_last_error()


@pytest.mark.parametrize("name", ["cfg4_bpx_L8", "bpx_L4", "pinv_bpx_L2"])
def test_common_controls_restrict_tile_sweeps(name):
    """Block-encoding circuits keep whole stretches under selector controls: a tile sweep whose gates
    all share controls outside its tile only launches the tiles where they are 1."""
    case = load_golden(name)
    n = case.n_qubits
    live = [g for g in case.gates if g.name != "I"]
    low = lower_gates(live, n)
    segs, _scalar, common, relabel = cabi.plan_describe(n, cabi.B2SV_C128, low, cabi.PlanOpts.make(mux=False), with_relabel=True)
    restricted = 0
    for (cls, mask, idx), (com, _n_ops) in zip(segs, common):
        if cls != cabi.SEG_TILE:
            assert com == 0
            continue
        assert com & mask == 0
        for gi in idx:
            ctrl = relabel[gi][3] if gi in relabel else int(low[gi]["ctrl_mask"])  # the SWAP frame may have renamed qubits
            assert ctrl & com == com, (name, gi)
        restricted += com != 0
    if n >= 17:
        assert restricted > 0

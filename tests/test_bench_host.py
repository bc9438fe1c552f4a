"""Host-side helpers of bench.py that need no GPU: the 'effective traffic' figure of SURVEY.md 8(d) and the timing of
the reference's Python projection predicate."""

import numpy as np

import bench
from tests.helpers import load_golden
from unitaria_b200 import lower_gates


def test_effective_bytes_counts_controls_and_half_diagonals():
    from unitaria_b200.fixtures import GateRecord as G

    n = 10
    gates = [
        G("H", (0,), (), None, 1.0),            # 2 * 16 * 2^10
        G("X", (1,), (0, 2), None, 1.0),        # two controls: 2 * 16 * 2^8
        G("Phase", (3,), (), 0.3, 1.0),         # diag(1, w): only target = 1 is touched: 2 * 16 * 2^9
        G("Rz", (4,), (5,), 0.3, 1.0),          # diag(w*, w) under one control: 2 * 16 * 2^9
    ]
    low = lower_gates(gates, n)
    want = 2 * 16 * (2**10 + 2**8 + 2**9 + 2**9)
    assert bench.effective_bytes(low, n, 16) == want
    assert bench.effective_bytes(low, n, 8) == want / 2


def test_effective_bytes_of_a_golden_exceeds_one_sweep_per_gate_only_without_controls():
    case = load_golden("laplace_n3")
    low = lower_gates(case.gates, case.n_qubits)
    eff = bench.effective_bytes(low, case.n_qubits, 16)
    assert 0 < eff <= len(low) * 2 * 16 * 2**case.n_qubits


def test_reference_projection_cost_reports_a_bounded_sample():
    case = load_golden("laplace_readme_n4")
    r = bench.reference_projection_cost(case.meta["subspace_out"], seconds=0.05)
    assert r["indices_sampled"] >= 2000 and r["us_per_index"] > 0
    assert np.isclose(r["enumerate_basis_seconds_at_full_size"], r["us_per_index"] * 1e-6 * 2 ** r["total_qubits"])

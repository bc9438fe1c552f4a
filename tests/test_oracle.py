"""The CPU oracle against the reference's golden vectors (CPU only, no GPU).

Every ``expected`` array in tests/golden/*.npz is the reference's own ``node.toarray()`` (its numpy
``compute()`` path); the check is the one ``ut.verify`` makes (verifier.py:115-126):
normalization * project(simulate(basis input)) == compute(e_i), here at 1e-12 instead of 1e-8.
"""

import numpy as np
import pytest

from oracle import c_oracle, np_oracle
from tests.helpers import column_error, expected_column, golden_inputs, golden_names, load_columns, load_golden, random_circuit, random_state

SMALL = golden_names(with_expected=True, max_qubits=16)


@pytest.mark.parametrize("name", SMALL)
def test_numpy_oracle_matches_reference_goldens(name):
    case = load_golden(name)
    n, spec_out, norm = case.n_qubits, case.meta["subspace_out"], case.meta["normalization"]
    limit = 6 if len(case.gates) > 5000 else 24
    for col, b in golden_inputs(case, limit=limit):
        got = np_oracle.node_simulate(case.gates, n, spec_out, norm, b)
        want = expected_column(case, col)
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= 1e-12 * scale * max(1, len(case.gates) / 100), (name, col)


@pytest.mark.parametrize("name", [n for n in SMALL if load_golden(n).n_qubits <= 14])
def test_c_oracle_matches_reference_goldens(name):
    case = load_golden(name)
    n, spec_out, norm = case.n_qubits, case.meta["subspace_out"], case.meta["normalization"]
    basis_out = np_oracle.enumerate_basis(spec_out)
    for col, b in golden_inputs(case, limit=6):
        psi = c_oracle.simulate_gates(case.gates, n, b)
        got = norm * psi[basis_out]
        want = expected_column(case, col)
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(got - want).max() <= 1e-12 * scale * max(1, len(case.gates) / 100), (name, col)


@pytest.mark.parametrize("name", ["laplace_n12", "laplace_n16", "laplace_n20", "gaussian_s3_t13", "bpx_L4", "bpx_L5", "cfg5_bits13"])
def test_c_oracle_matches_reference_columns_at_mid_size(name):
    """The oracle the GPU suite leans on at 15-23 qubits, pinned to the reference's node.compute(e_j) columns
    (tests/golden/make_golden.py --columns) -- also pins the rank -> basis-index map used to pick the inputs."""
    case = load_golden(name)
    cols, dim_out = load_columns(name)
    n, norm = case.n_qubits, case.meta["normalization"]
    basis_out = np_oracle.enumerate_basis(case.meta["subspace_out"])
    basis_in = np_oracle.enumerate_basis(case.meta["subspace_in"]) if case.meta["subspace_in"] is not None and n <= 20 else None
    assert len(basis_out) == dim_out
    for basis, idx, val in cols[:4]:
        if basis_in is not None:
            assert basis in basis_in
        got = norm * c_oracle.simulate_gates(case.gates, n, basis)[basis_out]
        assert column_error(got, idx, val) <= 1e-12 * max(1.0, float(np.abs(val).max())) * max(1, len(case.gates) / 1000), (name, basis)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11])
def test_c_oracle_equals_numpy_oracle_on_random_circuits(n):
    rng = np.random.default_rng(100 + n)
    circuit = random_circuit(n, 250, rng)
    psi0 = random_state(n, rng)
    a = np_oracle.simulate_gates(circuit.gates, n, psi0)
    b = c_oracle.simulate_gates(circuit.gates, n, psi0)
    assert np.abs(a - b).max() < 1e-13
    assert abs(np.linalg.norm(a) - 1) < 1e-12


def test_permutation_gates_are_exact_data_movement():
    # subspace.py:426-431 asserts `result[i] == 1` exactly after X/CNOT/Toffoli circuits
    rng = np.random.default_rng(7)
    circuit = random_circuit(9, 400, rng, kinds=["X", "CNOT", "Toffoli", "SWAP"])
    for b in (0, 5, 311):
        for sim in (np_oracle.simulate_gates, c_oracle.simulate_gates):
            out = sim(circuit.gates, 9, b)
            assert np.count_nonzero(out) == 1 and out[np.flatnonzero(out)[0]] == 1


def test_enumerate_basis_matches_goldens():
    for name in golden_names(with_expected=True, max_qubits=20):
        case = load_golden(name)
        if "basis_out" in case.arrays:
            assert np.array_equal(np_oracle.enumerate_basis(case.meta["subspace_out"]), case.arrays["basis_out"]), name
        if "basis_in" in case.arrays:
            assert np.array_equal(np_oracle.enumerate_basis(case.meta["subspace_in"]), case.arrays["basis_in"]), name


def test_test_basis_scalar_matches_vectorised():
    case = load_golden("bpx_L3")
    spec = case.meta["subspace_in"]
    mask = np_oracle.member_mask(spec)
    for i in range(0, mask.size, 37):
        assert np_oracle.test_basis(spec, i) == bool(mask[i])


def test_circuit_simulate_early_return_drops_global_phase():
    # circuit.py:54-61 (quirk Q1)
    import tequila as tq

    c = tq.gates.GlobalPhase(0.7)
    out = np_oracle.circuit_simulate(c, 2, 3)
    assert out.shape == (4,) and out[3] == 1
    vec = np.arange(4, dtype=complex)
    assert np_oracle.circuit_simulate(c, 2, vec) is vec

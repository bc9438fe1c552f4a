"""Host logic of the adapter on CPU, with the numpy test double standing in for the CUDA engine
(tests/fake_engine.py): the calls a unitaria user reaches through ``install()``."""

import numpy as np
import pytest

from oracle import np_oracle
from tests.helpers import load_golden
from unitaria_b200 import adapter


@pytest.fixture()
def fake_engine(monkeypatch):
    from tests.fake_engine import FakeStateVector

    monkeypatch.setattr(adapter, "StateVector", FakeStateVector)
    adapter.clear_caches()
    yield
    adapter.clear_caches()


class _Circuit:
    def __init__(self, gates):
        self.gates = gates


@pytest.mark.parametrize("name", ["laplace_n3", "laplace_n4", "bpx_L1", "increment4"])
def test_block_encoding_compute_follows_the_reference_formula(fake_engine, name):
    """nodes/block_encoding.py:45-66: embed the vector into the full register, simulate, normalization *
    project, and project once more -- vector and batch forms."""
    case = load_golden(name)
    if case.meta["is_vector"]:
        pytest.skip("needs a matrix-valued case")
    n, norm = case.n_qubits, case.meta["normalization"]
    s_in, s_out = case.meta["subspace_in"], case.meta["subspace_out"]
    b_in, b_out = np_oracle.enumerate_basis(s_in), np_oracle.enumerate_basis(s_out)
    if len(b_out) and b_out.max() >= len(b_out):
        pytest.skip("the reference's second projection is out of range for this subspace")
    rng = np.random.default_rng(5)
    batch = rng.normal(size=(3, len(b_in))) + 1j * rng.normal(size=(3, len(b_in)))

    def ref(vec):
        full = np.zeros(2**n, dtype=np.complex128)
        full[b_in] = vec
        out = norm * np_oracle.simulate_gates(case.gates, n, full)[b_out]
        return out[b_out]

    circuit = _Circuit(case.gates)
    got = adapter.block_encoding_compute(circuit, n, s_in, s_out, norm, batch[1])
    assert np.abs(got - ref(batch[1])).max() < 1e-12
    got = adapter.block_encoding_compute(circuit, n, s_in, s_out, norm, batch)
    assert got.shape == (3, len(b_out))
    assert np.abs(got - np.stack([ref(v) for v in batch])).max() < 1e-12


def test_block_encoding_compute_matches_the_columns_of_toarray(fake_engine):
    """Applied to unit vectors it reproduces the golden ``toarray()`` columns (what ut.verify compares)."""
    case = load_golden("laplace_n3")
    n, norm = case.n_qubits, case.meta["normalization"]
    s_in, s_out = case.meta["subspace_in"], case.meta["subspace_out"]
    expected = case.arrays["expected"]
    eye = np.eye(expected.shape[1], dtype=np.complex128)
    got = adapter.block_encoding_compute(_Circuit(case.gates), n, s_in, s_out, norm, eye)
    assert np.abs(got.T - expected).max() < 1e-12


@pytest.mark.parametrize("name", ["fourier4", "laplace_n3", "bpx_L2", "laplace_readme_n4"])
def test_block_encoding_compute_adjoint_follows_the_reference_formula(fake_engine, name):
    """nodes/block_encoding.py:68-80: project onto subspace_out, run the adjoint circuit on that vector as the
    state of the whole register, normalization * project onto subspace_in -- vector and batch forms.  The
    reference formula needs dimension_out == 2^n, so the output subspace here is the whole register."""
    from tests.helpers import dagger_gates

    case = load_golden(name)
    n, norm = case.n_qubits, case.meta["normalization"]
    s_in, s_out = case.meta["subspace_in"], [[[], []]] * n
    b_in = np_oracle.enumerate_basis(s_in)
    adjoint = _Circuit(dagger_gates(case.gates))
    rng = np.random.default_rng(6)
    batch = rng.normal(size=(3, 2**n)) + 1j * rng.normal(size=(3, 2**n))

    def ref(vec):
        return norm * np_oracle.simulate_gates(adjoint.gates, n, vec)[b_in]

    got = adapter.block_encoding_compute_adjoint(adjoint, n, s_in, s_out, norm, batch[2])
    assert got.shape == (len(b_in),) and np.abs(got - ref(batch[2])).max() < 1e-12
    got = adapter.block_encoding_compute_adjoint(adjoint, n, s_in, s_out, norm, batch)
    assert got.shape == (3, len(b_in))
    assert np.abs(got - np.stack([ref(v) for v in batch])).max() < 1e-12
    # a post-selected output subspace hands dimension_out < 2^n amplitudes to Circuit.simulate: an error upstream too
    if len(np_oracle.enumerate_basis(case.meta["subspace_out"])) != 2**n:
        with pytest.raises((ValueError, IndexError)):
            adapter.block_encoding_compute_adjoint(adjoint, n, s_in, case.meta["subspace_out"], norm, batch[0])

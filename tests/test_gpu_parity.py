"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference goldens.

Needs a B200 (``-m gpu``).  Tolerances are the north-star's: 1e-12 relative for complex128 storage,
1e-5 for complex64; permutation circuits and basis-state indexing bit-exact.
"""

import numpy as np
import pytest

from oracle import c_oracle, np_oracle
from tests.helpers import column_error, expected_column, golden_inputs, golden_names, load_columns, load_golden, random_circuit, random_state
from unitaria_b200 import PlanOpts, StateVector, SubspaceProgram, cabi, lower_gates
from unitaria_b200 import adapter

pytestmark = pytest.mark.gpu

# execution variants: every one must give the same answer
VARIANTS = {
    "fused": dict(),
    "unfused": dict(fuse=False),
    "tiles_only": dict(perm_runs=False),
    "tiles_plain_io": dict(perm_runs=False, plain_tile_io=True),
    "small_tiles": dict(tile_qubits=5, lookahead=4),
    "jit_sync": dict(jit="sync"),
    "jit_off": dict(jit="off"),
    "no_mux": dict(mux=False),
    "no_phases": dict(phases=False),
    "no_low6": dict(low6=False),
    "no_phases_no_mux": dict(phases=False, mux=False),
    "no_mono": dict(mono=False),
    "no_mono_tiles_only": dict(mono=False, perm_runs=False),
}


def tol(dtype, n_gates):
    base = 1e-12 if dtype == "complex128" else 1e-5
    return base * max(1.0, n_gates / 200.0)


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if cabi.device_count() == 0:
        pytest.fail("tests marked gpu need a CUDA device; the engine has no CPU fallback")
    yield
    adapter.clear_caches()


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", golden_names(with_expected=True, max_qubits=16))
def test_goldens_complex128(name, variant):
    """normalization * project(simulate(b)) == reference toarray()[:, col]  (verifier.py:115-126)."""
    case = load_golden(name)
    if variant != "fused" and len(case.gates) > 5000:
        pytest.skip("long circuit: fused variant only")
    n, norm = case.n_qubits, case.meta["normalization"]
    prog = SubspaceProgram(case.meta["subspace_out"])
    low = lower_gates(case.gates, n)
    opts = PlanOpts.make(**VARIANTS[variant])
    with StateVector(n, "complex128") as sv:
        for col, b in golden_inputs(case, limit=4 if len(case.gates) > 5000 else 12):
            sv.set_basis(b)
            sv.apply_lowered(low, opts)
            got = sv.project(prog, scale=norm)
            want = expected_column(case, col)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(got - want).max() <= tol("complex128", len(case.gates)) * scale, (name, variant, col)
            n2 = sv.project_norm2(prog)
            assert abs(abs(norm) * np.sqrt(n2) - np.linalg.norm(want)) <= 1e-11 * max(1.0, np.linalg.norm(want))


@pytest.mark.parametrize("name", golden_names(with_expected=True, max_qubits=16))
def test_goldens_complex64(name):
    case = load_golden(name)
    n, norm = case.n_qubits, case.meta["normalization"]
    prog = SubspaceProgram(case.meta["subspace_out"])
    low = lower_gates(case.gates, n)
    with StateVector(n, "complex64") as sv:
        for col, b in golden_inputs(case, limit=4):
            sv.set_basis(b)
            sv.apply_lowered(low)
            got = sv.project(prog, scale=norm)
            want = expected_column(case, col)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(got - want).max() <= tol("complex64", len(case.gates)) * scale, (name, col)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("n", [1, 2, 3, 4, 6, 9, 12, 13, 15])
def test_random_circuits_full_state(n, variant):
    rng = np.random.default_rng(1000 + n)
    circuit = random_circuit(n, 300, rng, max_controls=min(4, max(0, n - 2)))
    psi0 = random_state(n, rng)
    want = c_oracle.simulate_gates(circuit.gates, n, psi0)
    with StateVector(n, "complex128") as sv:
        sv.upload(psi0)
        sv.apply(circuit.gates, PlanOpts.make(**VARIANTS[variant]))
        got = sv.download()
    assert np.abs(got - want).max() <= 1e-12, (n, variant)
    with StateVector(n, "complex64") as sv:
        sv.upload(psi0)
        sv.apply(circuit.gates, PlanOpts.make(**VARIANTS[variant]))
        got = sv.download()
    assert np.abs(got - want).max() <= 1e-5, (n, variant)


@pytest.mark.parametrize("variant", ["fused", "tiles_only", "unfused", "jit_sync", "jit_off"])
@pytest.mark.parametrize("n", [13, 14, 17, 21])
def test_permutation_circuits_are_bit_exact(n, variant):
    """X / CNOT / Toffoli / SWAP circuits move amplitudes without arithmetic (subspace.py:426-431)."""
    rng = np.random.default_rng(n)
    circuit = random_circuit(n, 500, rng, kinds=["X", "CNOT", "Toffoli", "SWAP"], max_controls=5)
    psi0 = random_state(n, rng)
    want = c_oracle.simulate_gates(circuit.gates, n, psi0)
    for dtype, ref in (("complex128", want), ("complex64", want.astype(np.complex64).astype(np.complex128))):
        with StateVector(n, dtype) as sv:
            sv.upload(psi0)
            sv.apply(circuit.gates, PlanOpts.make(**VARIANTS[variant]))
            got = sv.download()
        assert np.array_equal(got, ref), (n, variant, dtype)
    with StateVector(n) as sv:
        sv.set_basis(5)
        sv.apply(circuit.gates, PlanOpts.make(**VARIANTS[variant]))
        got = sv.download()
    assert np.count_nonzero(got) == 1 and got[np.flatnonzero(got)[0]] == 1


def _perm_heavy(n, m, rng, max_controls=4):
    kinds = ["X"] * 8 + ["CNOT"] * 3 + ["Toffoli"] * 3 + ["SWAP"] * 3 + ["Phase", "GlobalPhase", "Z", "Zp", "Rz"] * 2 + ["Ry", "H"]
    return random_circuit(n, m, rng, kinds=kinds, max_controls=max_controls)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("n,seed,tile_qubits", [(7, 0, 5), (9, 1, 6), (12, 2, 8), (14, 3, 0), (15, 4, 12), (16, 5, 13), (17, 6, 0), (20, 7, 0)])
def test_monomial_phases_random(n, seed, tile_qubits, dtype):
    """Runs of X / SWAP / phase gates inside tile sweeps (kernels.cuh: tile_apply_mono) against the oracle:
    the same circuits the CPU suite pushes through the host emulator (tests/test_tile_emul.py)."""
    rng = np.random.default_rng(seed)
    circuit = _perm_heavy(n, 600, rng)
    psi0 = random_state(n, rng)
    want = c_oracle.simulate_gates(circuit.gates, n, psi0)
    with StateVector(n, dtype) as sv:
        for mono in (True, False):
            sv.upload(psi0)
            sv.apply(circuit.gates, PlanOpts.make(tile_qubits=tile_qubits, mono=mono, perm_runs=False))
            got = sv.download()
            assert np.abs(got - want).max() < tol(dtype, 600), (mono,)


@pytest.mark.parametrize("dtype,n", [("complex128", 15), ("complex64", 15), ("complex128", 22)])
def test_monomial_long_runs_are_bit_exact_and_batched(dtype, n):
    """Several hundred X-family gates confined to ten qubits ride in ONE tile sweep: more ops than one staged
    batch, four rounds per phase for complex64 tiles.  Pure permutations must be exact data movement."""
    rng = np.random.default_rng(7)
    circuit = random_circuit(n, 900, rng, kinds=["X", "CNOT", "Toffoli", "SWAP"], max_controls=3)
    keep = [g for g in circuit.gates if all(t < 10 for t in g.target)]
    psi0 = random_state(n, rng)
    if dtype == "complex64":
        psi0 = psi0.astype(np.complex64).astype(np.complex128)
    want = c_oracle.simulate_gates(keep, n, psi0)
    with StateVector(n, dtype) as sv:
        sv.upload(psi0)
        sv.stats_reset()
        sv.apply(keep, PlanOpts.make(perm_runs=False))
        got = sv.download()
        launches = sv.stats()["launches_by_class"]
    assert np.array_equal(got, want)
    assert 1 <= launches["tile"] <= 16, launches


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("jit", ["sync", "off"])
@pytest.mark.parametrize("name", ["laplace_n12", "laplace_n16", "laplace_n20", "gaussian_s3_t13", "cfg5_bits13"])
def test_basis_inputs_through_the_live_window(name, jit, dtype):
    """set_basis + [one tile sweep] + [permutation sweep]: only the tile that owns the index is written and the
    permutation sweep takes zeros for the rest (b2sv::window).  Full-state parity for basis indices in
    different tiles, then readers that need the whole register right after the first sweep."""
    case = load_golden(name)
    n = case.n_qubits
    live = [g for g in case.gates if g.name != "I"]  # lowering drops the I pads: keep plan indices aligned
    low = lower_gates(live, n)
    opts = PlanOpts.make(jit=jit)
    with StateVector(n, dtype) as sv:
        segs, _ = cabi.plan_describe(n, cabi.B2SV_C128 if dtype == "complex128" else cabi.B2SV_C64, low, opts)
        classes = [c for c, _m, _i in segs]
        assert classes[0] in (cabi.SEG_TILE, cabi.SEG_LOW6) and cabi.SEG_PERM in classes[1:2], classes  # the shape this test is about
        for index in (0, 1, 5, (1 << n) - 1, 1 << (n - 1), (1 << (n - 2)) + 77):
            psi0 = np.zeros(2**n, dtype=np.complex128)
            psi0[index] = 1
            want = c_oracle.simulate_gates(case.gates, n, psi0)
            # dirty the buffers first: the window must not read stale amplitudes
            sv.upload(random_state(n, np.random.default_rng(index)))
            sv.apply_lowered(low, opts)
            sv.set_basis(index)
            sv.apply_lowered(low, opts)
            got = sv.download()
            assert np.abs(got - want).max() < tol(dtype, len(case.gates)), index
        # a prefix that ends right after the first sweep: the next reader materialises the implied zeros
        first = [live[i] for i in segs[0][2]]
        sv.upload(random_state(n, np.random.default_rng(1)))
        sv.set_basis(3)
        sv.apply(first, opts)
        psi0 = np.zeros(2**n, dtype=np.complex128)
        psi0[3] = 1
        assert np.abs(sv.download() - c_oracle.simulate_gates(first, n, psi0)).max() < tol(dtype, len(first))


@pytest.mark.parametrize("name", ["laplace_n12", "laplace_n16", "laplace_n20", "gaussian_s3_t13", "bpx_L4", "bpx_L5", "cfg5_bits13"])
def test_mid_size_configs_against_c_oracle(name):
    """The BASELINE.json circuit families at sizes the CPU oracle finishes in seconds: full state."""
    case = load_golden(name)
    n = case.n_qubits
    rng = np.random.default_rng(3)
    psi0 = random_state(n, rng)
    want = c_oracle.simulate_gates(case.gates, n, psi0)
    low = lower_gates(case.gates, n)
    with StateVector(n, "complex128") as sv:
        sv.upload(psi0)
        sv.apply_lowered(low)
        got = sv.download()
        assert np.abs(got - want).max() <= 1e-12, name
        for variant in ("tiles_only", "jit_sync", "jit_off"):
            sv.upload(psi0)
            sv.apply_lowered(low, PlanOpts.make(**VARIANTS[variant]))
            got2 = sv.download()
            assert np.abs(got2 - want).max() <= 1e-12, (name, variant)
    with StateVector(n, "complex64") as sv:
        sv.upload(psi0)
        sv.apply_lowered(low)
        got = sv.download()
        assert np.abs(got - want).max() <= 1e-5, name


@pytest.mark.parametrize("dtype,tol", [("complex128", 1e-12), ("complex64", 1e-5)])
@pytest.mark.parametrize("name", ["laplace_n12", "laplace_n16", "laplace_n20", "gaussian_s3_t13", "bpx_L4", "bpx_L5", "cfg5_bits13"])
def test_mid_size_columns_match_reference_compute(name, dtype, tol):
    """The mid-size circuit families against the reference's own node.compute(e_j) columns (verifier.py:115-126)."""
    case = load_golden(name)
    cols, dim_out = load_columns(name)
    assert cols, "run tests/golden/make_golden.py --columns"
    n, norm = case.n_qubits, case.meta["normalization"]
    prog = SubspaceProgram(case.meta["subspace_out"])
    assert prog.dimension == dim_out
    low = lower_gates(case.gates, n)
    with StateVector(n, dtype) as sv:
        for basis, idx, val in cols:
            sv.set_basis(basis)
            sv.apply_lowered(low)
            got = sv.project(prog, scale=norm)
            scale = max(1.0, float(np.abs(val).max()) if len(val) else 1.0)
            assert column_error(got, idx, val) <= tol * scale, (name, dtype, basis)


@pytest.mark.parametrize("name", ["laplace_readme_n8", "bpx_L3", "pinv_bpx_L2", "cfg5_bits13", "laplace_n16"])
@pytest.mark.parametrize("which", ["subspace_in", "subspace_out"])
def test_device_enumerate_basis_and_projection_order(name, which):
    """Device predicate + rank == Subspace.enumerate_basis order (subspace.py:325-329), bit-exact."""
    case = load_golden(name)
    spec = case.meta[which]
    prog = SubspaceProgram(spec)
    if prog.total_qubits > 22:
        pytest.skip("oracle enumeration too large")
    want = np_oracle.enumerate_basis(spec)
    n = case.n_qubits
    rng = np.random.default_rng(11)
    psi = random_state(n, rng)
    with StateVector(n) as sv:
        got = sv.enumerate_basis(prog)
        assert np.array_equal(got, want)
        sv.upload(psi)
        assert np.array_equal(sv.project(prog), psi[want])  # pure gather: bit-exact
        assert np.abs(sv.project(prog, scale=2.5 - 1j) - (2.5 - 1j) * psi[want]).max() < 1e-15
        assert abs(sv.project_norm2(prog) - float(np.sum(np.abs(psi[want]) ** 2))) < 1e-13


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("name", ["laplace_readme_n8", "bpx_L3", "pinv_bpx_L2", "cfg5_bits13", "laplace_n16"])
def test_embed_is_the_inverse_of_project(name, dtype):
    """b2sv_embed: the members take the vector in enumerate_basis order, everything else is zero
    (BlockEncoding.compute's ``full_input[basis] = input``, nodes/block_encoding.py:49-52)."""
    case = load_golden(name)
    spec = case.meta["subspace_in"]
    prog = SubspaceProgram(spec)
    if prog.total_qubits > 22:
        pytest.skip("oracle enumeration too large")
    basis = np_oracle.enumerate_basis(spec)
    n = case.n_qubits
    rng = np.random.default_rng(12)
    vec = rng.normal(size=len(basis)) + 1j * rng.normal(size=len(basis))
    if dtype == "complex64":
        vec = vec.astype(np.complex64).astype(np.complex128)
    want = np.zeros(2**n, dtype=np.complex128)
    want[basis] = vec
    with StateVector(n, dtype) as sv:
        sv.set_basis(3)  # a pending basis state must not survive the embedding
        sv.embed(prog, vec)
        assert np.array_equal(sv.download(), want)
        assert np.array_equal(sv.project(prog), vec)
        with pytest.raises(ValueError):
            sv.embed(prog, vec[:-1])


@pytest.mark.parametrize("name", ["laplace_n8", "bpx_L3", "gaussian_s3_t8", "laplace_readme_n8"])
def test_block_encoding_compute_device_resident(name):
    """adapter.block_encoding_compute == the reference's BlockEncoding.compute formula evaluated with the
    oracle: embed, simulate, normalization * project, project again (block_encoding.py:45-66)."""
    case = load_golden(name)
    if case.meta["is_vector"]:
        pytest.skip("needs a matrix-valued case")
    n, norm = case.n_qubits, case.meta["normalization"]
    s_in, s_out = case.meta["subspace_in"], case.meta["subspace_out"]
    b_in, b_out = np_oracle.enumerate_basis(s_in), np_oracle.enumerate_basis(s_out)
    if len(b_out) and b_out.max() >= len(b_out):
        pytest.skip("the reference's second projection is out of range for this subspace")
    rng = np.random.default_rng(13)
    batch = rng.normal(size=(3, len(b_in))) + 1j * rng.normal(size=(3, len(b_in)))

    class Circuit:
        gates = case.gates

    def ref(vec):
        full = np.zeros(2**n, dtype=np.complex128)
        full[b_in] = vec
        out = norm * np_oracle.simulate_gates(case.gates, n, full)[b_out]
        return out[b_out]

    got = adapter.block_encoding_compute(Circuit, n, s_in, s_out, norm, batch[0])
    assert np.abs(got - ref(batch[0])).max() < tol("complex128", len(case.gates))
    got = adapter.block_encoding_compute(Circuit, n, s_in, s_out, norm, batch)
    want = np.stack([ref(v) for v in batch])
    assert got.shape == want.shape and np.abs(got - want).max() < tol("complex128", len(case.gates))


@pytest.mark.parametrize("name", ["fourier4", "laplace_n8", "bpx_L3", "laplace_readme_n8"])
def test_block_encoding_compute_adjoint_device_resident(name):
    """adapter.block_encoding_compute_adjoint == the reference's BlockEncoding.compute_adjoint formula evaluated
    with the oracle (block_encoding.py:68-80): the adjoint circuit on the projected input as the state of the whole
    register, normalization * projection onto subspace_in gathered on the device."""
    from tests.helpers import dagger_gates

    case = load_golden(name)
    n, norm = case.n_qubits, case.meta["normalization"]
    s_in, s_out = case.meta["subspace_in"], [[[], []]] * n  # the formula needs dimension_out == 2^n
    b_in = np_oracle.enumerate_basis(s_in)
    rng = np.random.default_rng(17)
    batch = (rng.normal(size=(2, 2**n)) + 1j * rng.normal(size=(2, 2**n))) / np.sqrt(2.0 ** (n + 1))

    class Adjoint:
        gates = dagger_gates(case.gates)

    def ref(vec):
        return norm * c_oracle.simulate_gates(Adjoint.gates, n, vec)[b_in]

    got = adapter.block_encoding_compute_adjoint(Adjoint, n, s_in, s_out, norm, batch[1])
    assert got.shape == (len(b_in),) and np.abs(got - ref(batch[1])).max() < tol("complex128", len(case.gates))
    got = adapter.block_encoding_compute_adjoint(Adjoint, n, s_in, s_out, norm, batch)
    want = np.stack([ref(v) for v in batch])
    assert got.shape == want.shape and np.abs(got - want).max() < tol("complex128", len(case.gates))
    adapter.clear_caches()


def test_adapter_simulate_circuit_contract():
    """Circuit.simulate's contract (circuit.py:42-70) incl. quirks Q1-Q3, Q7 of SURVEY.md App. D."""
    import tequila as tq

    c = tq.gates.Ry(1.854590436003224, 0)  # README.md:20-24: encodes (3,4)/5
    out = adapter.simulate_circuit(c, 1, 0, backend="qulacs")
    assert out.dtype == np.complex128 and out.shape == (2,)
    assert np.allclose(out, [0.6, 0.8], atol=1e-15)
    # np.integer input, dense input, register wider than the used qubits
    c2 = tq.gates.X(0) + tq.gates.CNOT(0, 2)
    assert adapter.simulate_circuit(c2, 4, np.int64(0))[5] == 1
    vec = np.zeros(16, dtype=complex)
    vec[2] = 1j
    assert adapter.simulate_circuit(c2, 4, vec)[2 | 1 | 4] == 1j
    # no qubit touched: returns input / one-hot and DROPS the global phase
    g = tq.gates.GlobalPhase(0.9)
    assert adapter.simulate_circuit(g, 2, 1)[1] == 1
    assert adapter.simulate_circuit(g, 2, vec[:4]) is not None
    # controlled global phase is NOT dropped
    cg = tq.gates.X(0) + tq.gates.GlobalPhase(0.9).add_controls([0])
    assert abs(adapter.simulate_circuit(cg, 1, 0)[1] - np.exp(0.9j)) < 1e-15
    # uncontrolled global phase next to real gates is applied
    cg2 = tq.gates.GlobalPhase(0.4) + tq.gates.X(1)
    assert abs(adapter.simulate_circuit(cg2, 2, 0)[2] - np.exp(0.4j)) < 1e-15


def test_adapter_projected_and_norm_match_oracle():
    case = load_golden("laplace_n3")  # BASELINE config 1

    class C:
        gates = case.gates

    n, norm, spec = case.n_qubits, case.meta["normalization"], case.meta["subspace_out"]
    for col, b in golden_inputs(case):
        got = adapter.simulate_projected(C, n, spec, norm, b)
        assert np.abs(got - expected_column(case, col)).max() < 1e-13
        nrm = adapter.simulate_norm(C, n, spec, norm, b)
        assert abs(nrm - np.linalg.norm(expected_column(case, col))) < 1e-13
        full = adapter.simulate_circuit(C, n, b)
        assert np.abs(full - np_oracle.simulate_gates(case.gates, n, b)).max() < 1e-14


def test_dagger_roundtrip_restores_state():
    """encode -> inverse round trip at a size no oracle array is stored for."""
    import tequila as tq

    n = 22
    rng = np.random.default_rng(4)
    circuit = random_circuit(n, 200, rng, max_controls=3)
    both = tq.QCircuit(circuit.gates) + circuit.dagger()
    psi0 = random_state(n, rng)
    with StateVector(n) as sv:
        sv.upload(psi0)
        sv.apply(both.gates)
        got = sv.download()
    assert np.abs(got - psi0).max() < 1e-12


def test_bad_arguments_are_rejected():
    with StateVector(3) as sv:
        with pytest.raises(cabi.B2svError):
            sv.set_basis(8)
        with pytest.raises(ValueError):
            sv.upload(np.zeros(4, dtype=complex))
        bad = np.zeros(1, dtype=cabi.GATE_DTYPE)
        bad["op"][0] = cabi.OP_X
        bad["t0"][0] = 5
        with pytest.raises(cabi.B2svError):
            sv.apply_lowered(bad)
        with pytest.raises(cabi.B2svError):
            sv.project_norm2("#####")  # a 5-qubit subspace on a 3-qubit register


def test_sampling_mode_matches_exact_norm_statistically():
    """BackendEstimator's seam (estimator/backend_estimator.py:39-51): post-selected shot counts.  The
    engine draws the two binomial counts from exact device reductions; with N shots the estimate must
    sit within a few standard errors of the exact norm."""
    case = load_golden("constant_vector5")  # ConstantVector(5 entries) on 3 qubits, subspace from_dim(5)

    class C:
        gates = case.gates

    n, norm, spec = case.n_qubits, case.meta["normalization"], case.meta["subspace_out"]
    exact = float(np.linalg.norm(case.arrays["expected"]))
    rng = np.random.default_rng(7)
    samples = 200_000
    est = [adapter.estimate_norm_sampled(C, n, spec, norm, samples, rng=rng) for _ in range(8)]
    p = (exact / norm) ** 2
    stderr = norm * np.sqrt(p * (1 - p) / samples) / (2 * np.sqrt(p)) if 0 < p < 1 else 1e-9
    assert abs(np.mean(est) - exact) < 6 * stderr / np.sqrt(len(est)) + 1e-9
    assert all(abs(e - exact) < 8 * stderr + 1e-9 for e in est)


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
def test_fill_random_matches_its_host_restatement(dtype):
    """b2sv_fill_random (device-side seeded dense vector for benchmarks / full-size property tests) == the numpy
    restatement of the same counter hash, including the support window and a shard's index base."""
    from unitaria_b200.engine import random_state_hash

    n = 12
    with StateVector(n, dtype) as sv:
        for seed, bits, base in [(0, None, 0), (7, 9, 0), (3, 14, 3 << n)]:
            sv.set_basis(5)  # pending state must not survive
            sv.fill_random(seed=seed, support_bits=bits, index_base=base)
            got = sv.download()
            want = random_state_hash(n, seed, bits if bits is not None else n, base)
            tol = 0 if dtype == "complex128" else 1e-7
            assert np.abs(got - want).max() <= tol, (seed, bits, base)
            if bits is None:
                assert abs(np.linalg.norm(got) - 1.0) < 0.05


def test_known_zero_state_is_lazy_and_correct():
    """b2sv_set_zero writes nothing: gates are skipped, the norm is 0 without a launch, readers get zeros."""
    import tequila as tq

    n = 14
    c = tq.gates.H(0) + tq.gates.CNOT(0, 5) + tq.gates.Ry(0.3, 9)
    with StateVector(n) as sv:
        sv.upload(random_state(n, np.random.default_rng(0)))  # dirty buffer
        sv.set_zero()
        sv.stats_reset()
        sv.apply(c.gates)
        assert sv.project_norm2("#" * n) == 0.0
        assert sv.stats()["launches"] == 0
        assert not np.any(sv.download())
        sv.set_basis(3)
        sv.apply(c.gates)
        assert abs(sv.project_norm2("#" * n) - 1.0) < 1e-14


def test_two_devices_in_one_process():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device -- a second StateVector on another
    device of the same process must run the 64 KiB tile sweeps too."""
    if cabi.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = load_golden("laplace_n12")
    psi0 = random_state(case.n_qubits, np.random.default_rng(2))
    want = c_oracle.simulate_gates(case.gates, case.n_qubits, psi0)
    for device in (0, 1, 0):
        with StateVector(case.n_qubits, device=device) as sv:
            sv.upload(psi0)
            sv.apply(case.gates)
            assert np.abs(sv.download() - want).max() < 1e-12, device


def test_device_resident_chaining_and_pipelined_columns():
    """results that stay on the GPU (torch tensor views), a CUDA tensor as the initial state, and
    the pipelined matrix-case loop of Node.simulate (node.py:381-388) -- same numbers as the host path."""
    import torch

    case = load_golden("laplace_n8")

    class C:
        gates = case.gates

    n, norm, spec = case.n_qubits, case.meta["normalization"], case.meta["subspace_out"]
    inputs = [b for _c, b in golden_inputs(case, limit=9)]
    host = [adapter.simulate_projected(C, n, spec, norm, b).copy() for b in inputs]
    cols = adapter.simulate_columns(C, n, spec, norm, inputs)
    for a, b in zip(host, cols):
        assert np.array_equal(a, b)
    recycled = adapter.simulate_columns(C, n, spec, norm, inputs, keep=False)
    assert np.array_equal(recycled[-1], host[-1]) and np.array_equal(recycled[-2], host[-2])
    dev = adapter.simulate_projected(C, n, spec, norm, inputs[3], out="device")
    assert dev.is_cuda and dev.dtype == torch.complex128 and np.array_equal(dev.cpu().numpy(), host[3])
    # chain: feed a device-resident full state back in (dense initial state that never touches the host)
    psi0 = random_state(n, np.random.default_rng(3))
    want = adapter.simulate_projected(C, n, spec, norm, psi0).copy()
    got = adapter.simulate_projected(C, n, spec, norm, torch.as_tensor(psi0, device="cuda:0"), out="device")
    assert np.abs(got.cpu().numpy() - want).max() < 1e-14


@pytest.mark.parametrize("dtype", ["complex128", "complex64"])
@pytest.mark.parametrize("name", ["laplace_n16", "laplace_n20", "cfg5_bits13"])
def test_sparse_projection_equals_dense(name, dtype):
    """b2sv_project_gather_sparse lists exactly the non-zeros of b2sv_project_gather (bit for bit), project() assembles
    the same dense array from them, and a dense state makes the compact form step aside (nnz >= cap -> None)."""
    case = load_golden(name)
    n, norm = case.n_qubits, case.meta["normalization"]
    prog = SubspaceProgram(case.meta["subspace_out"])
    low = lower_gates(case.gates, n)
    with StateVector(n, dtype) as sv:
        for b in (0, 5):
            sv.set_basis(b)
            sv.apply_lowered(low)
            saved = StateVector.SPARSE_MIN_DIM
            StateVector.SPARSE_MIN_DIM = 1 << 62
            try:
                dense = sv.project(prog, scale=norm).copy()
                assert sv.last_d2h_bytes == 16 * prog.dimension
            finally:
                StateVector.SPARSE_MIN_DIM = saved
            pairs = sv.project_sparse(prog, scale=norm)
            assert pairs is not None, "a basis-state column of these block encodings has few non-zeros"
            idx, val = pairs
            order = np.argsort(idx)
            want = np.flatnonzero(dense)
            assert np.array_equal(idx[order].astype(np.int64), want)
            assert np.array_equal(val[order], dense[want])
            StateVector.SPARSE_MIN_DIM = 1
            try:
                again = sv.project(prog, scale=norm)
                assert sv.last_d2h_bytes == 8 + 24 * len(want)
            finally:
                StateVector.SPARSE_MIN_DIM = saved
            assert np.array_equal(again, dense)
        rng = np.random.default_rng(11)
        sv.upload(random_state(n, rng))
        assert sv.project_sparse("#" * n) is None
        StateVector.SPARSE_MIN_DIM = 1
        try:
            full = sv.project("#" * n)
            assert sv.last_d2h_bytes == 16 * 2**n and np.array_equal(full, sv.download())
            sv.project("#" * n)  # (the handle skips the compact attempt for a while after a dense result)
            assert sv.last_d2h_bytes == 16 * 2**n
        finally:
            StateVector.SPARSE_MIN_DIM = saved

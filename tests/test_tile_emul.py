"""Tile-op programs (host side of the tile sweeps), checked on CPU.

``tests/emul/tile_emul.cu`` is a host-only interpreter of the op programs that
``unitaria_b200/csrc/tileops.hpp`` builds for the CUDA tile kernels: tile-local positions, fixed-position
enumeration, register phases, monomial phases (round masks, shared controls), batch padding.  The
planner and op builders are the product's, included unchanged; the result is compared with the CPU
oracle.  The device code that interprets the same arrays is covered by the ``-m gpu`` parity tests.
"""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import np_oracle
from tests.helpers import load_golden, random_circuit, random_state
from unitaria_b200 import cabi
from unitaria_b200.lowering import lower_gates

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "tile_emul.cu")
LIB = os.path.join(HERE, "emul", "libtile_emul.so")
CSRC = os.path.join(os.path.dirname(HERE), "unitaria_b200", "csrc")
INFO = ["tile", "perm", "gate1q", "low6", "mono_phases", "mono_members", "reg_phases", "reg_members", "max_ops", "nops", "mono_rounds_max", "mono_gates"]


def _build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("kernels.cuh", "planner.hpp", "tileops.hpp")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-o", LIB, SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return LIB


@pytest.fixture(scope="module")
def emul():
    lib = ctypes.CDLL(_build())
    lib.emul_apply.restype = ctypes.c_int
    lib.emul_apply.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    def run(gates, n, psi0, dtype=cabi.B2SV_C128, **kw):
        live = [g for g in gates if g.name != "I"]
        low = np.ascontiguousarray(lower_gates(live, n), dtype=cabi.GATE_DTYPE)
        opts = cabi.PlanOpts.make(**kw)
        amps = np.ascontiguousarray(psi0, dtype=np.complex128).copy()
        info = np.zeros(len(INFO), dtype=np.int64)
        rc = lib.emul_apply(n, dtype, low.ctypes.data, len(low), ctypes.byref(opts), amps.ctypes.data, info.ctypes.data)
        assert rc == 0, f"emulator rejected the program: rc={rc}"
        return amps, dict(zip(INFO, (int(x) for x in info)))

    return run


def perm_heavy_circuit(n, m, rng, max_controls=4):
    """Mostly X / SWAP / phase gates with a few rotations: what ripple-carry arithmetic under selector
    controls looks like (circuits/arithmetic.py), so that long monomial runs form."""
    kinds = ["X"] * 8 + ["CNOT"] * 3 + ["Toffoli"] * 3 + ["SWAP"] * 3 + ["Phase", "GlobalPhase", "Z", "Zp", "Rz"] * 2 + ["Ry", "H"]
    return random_circuit(n, m, rng, kinds=kinds, max_controls=max_controls)


@pytest.mark.parametrize("n,seed,tile_qubits", [(7, 0, 5), (9, 1, 6), (12, 2, 8), (13, 6, 11), (14, 3, 0), (15, 4, 12), (15, 5, 13)])
@pytest.mark.parametrize("mono", [True, False])
def test_random_permutation_heavy(emul, n, seed, tile_qubits, mono):
    rng = np.random.default_rng(seed)
    circuit = perm_heavy_circuit(n, 600, rng)
    psi0 = random_state(n, rng)
    want = np_oracle.simulate_gates([g for g in circuit.gates if g.name != "I"], n, psi0)
    got, info = emul(circuit.gates, n, psi0, tile_qubits=tile_qubits, mono=mono, perm_runs=False)
    assert np.abs(got - want).max() < 1e-12
    k = tile_qubits or 12
    if mono and k >= 11:  # a thread needs at least eight amplitudes of its own
        assert info["mono_phases"] > 0, info
    else:
        assert info["mono_phases"] == 0


@pytest.mark.parametrize("n,seed", [(8, 10), (13, 11), (14, 12)])
@pytest.mark.parametrize("dtype", [cabi.B2SV_C128, cabi.B2SV_C64])
def test_random_whole_alphabet(emul, n, seed, dtype):
    rng = np.random.default_rng(seed)
    circuit = random_circuit(n, 500, rng)
    psi0 = random_state(n, rng)
    want = np_oracle.simulate_gates([g for g in circuit.gates if g.name != "I"], n, psi0)
    got, _ = emul(circuit.gates, n, psi0, dtype=dtype)
    assert np.abs(got - want).max() < 1e-12


def test_long_runs_are_batched_and_rounds_are_used(emul):
    """A run of several hundred X-family gates on <= 11 qubits inside one sweep: more ops than one staged
    batch holds, phases never straddle a batch boundary (the emulator rejects that), 13-qubit tiles need
    four rounds per monomial phase."""
    n = 15
    rng = np.random.default_rng(7)
    circuit = random_circuit(n, 900, rng, kinds=["X", "CNOT", "Toffoli", "SWAP", "Phase"], max_controls=3)
    # confine the moved qubits to ten of them so that the run fits one tile sweep
    keep = [g for g in circuit.gates if all(t < 10 for t in g.target)]
    psi0 = random_state(n, rng)
    want = np_oracle.simulate_gates(keep, n, psi0)
    for dtype, rounds in ((cabi.B2SV_C128, 2), (cabi.B2SV_C64, 4)):
        # a handful of headers: the walks were done on the host
        got, info = emul(keep, n, psi0, dtype=dtype, perm_runs=False)
        assert np.abs(got - want).max() < 1e-12
        assert info["mono_phases"] > 0 and info["max_ops"] < 112 and info["mono_rounds_max"] == rounds, info
    # dense gates on ten qubits cannot be grouped like that: one sweep, more ops than one staged batch holds
    dense = [g for g in random_circuit(n, 700, rng, kinds=["Ry", "H", "Rz", "CNOT"], max_controls=1).gates if all(t < 10 for t in g.target)]
    want = np_oracle.simulate_gates(dense, n, psi0)
    got, info = emul(dense, n, psi0, perm_runs=False, mux=False)
    assert np.abs(got - want).max() < 1e-12
    assert info["max_ops"] > 112, info


@pytest.mark.parametrize("name", ["bpx_L3", "bpx_L4", "pinv_bpx_L2", "laplace_n12", "gaussian_s3_t8", "cfg5_bits5", "qsvt_inc2_cheb", "sinv_rhs_bpx_L2"])
def test_goldens_through_the_tile_programs(emul, name):
    case = load_golden(name)
    n = case.n_qubits
    rng = np.random.default_rng(3)
    psi0 = random_state(n, rng)
    live = [g for g in case.gates if g.name != "I"]
    want = np_oracle.simulate_gates(live, n, psi0)
    got, info = emul(case.gates, n, psi0)
    assert np.abs(got - want).max() < 1e-12, info


@pytest.mark.parametrize("block", range(4))
def test_fuzz_random_shapes(emul, block):
    """Random sizes, tile widths, control counts, look-ahead windows and gate mixes (with and without targets
    confined to a few qubits, which is what makes long compact runs): the builder must never produce a
    program the interpreter rejects, and the result must match the oracle."""
    mixes = [
        ["X"] * 8 + ["CNOT"] * 3 + ["Toffoli"] * 3 + ["SWAP"] * 3 + ["Phase", "GlobalPhase", "Z", "Zp", "Rz"] * 2 + ["Ry", "H"],
        ["X", "CNOT", "Toffoli", "SWAP", "Phase", "Rz", "GlobalPhase", "Z"],
        ["X", "SWAP", "Rz", "Phase"] * 3 + ["Ry"],
        None,
    ]
    for seed in range(10 * block, 10 * block + 10):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.integers(11, 16))
        tile_qubits = int(rng.choice([0, 11, 12, 13])) if n >= 13 else 0
        dtype = cabi.B2SV_C64 if rng.integers(0, 2) else cabi.B2SV_C128
        circuit = random_circuit(n, int(rng.integers(50, 500)), rng, kinds=mixes[int(rng.integers(0, 4))], max_controls=int(rng.integers(1, 7)))
        gates = [g for g in circuit.gates if g.name != "I"]
        if rng.integers(0, 2):
            limit = int(rng.integers(6, n))
            gates = [g for g in gates if all(t < limit for t in g.target)]
        if not gates:
            continue
        psi0 = random_state(n, rng)
        want = np_oracle.simulate_gates(gates, n, psi0)
        got, _ = emul(gates, n, psi0, dtype=dtype, tile_qubits=tile_qubits, perm_runs=bool(rng.integers(0, 2)), lookahead=int(rng.choice([0, 4, 32])))
        assert np.abs(got - want).max() < 1e-11, seed


def test_small_multiplexors_ride_in_dense_blocks(emul):
    """A uniformly controlled rotation tree on three qubits (circuits/multiplexed_rot.py: Ry / CNOT ladders, the closing
    sweep of the Gaussian-convolution circuit) is ONE 8x8 matrix: the multiplexor fusion's table op (index bits 1, 2 ->
    target 0) must be multiplied into the dense block on {0, 1, 2}, not run as a pass of its own -- and the result must
    not change."""
    from unitaria_b200.fixtures import GateRecord as G

    n = 16
    rng = np.random.default_rng(11)
    gates = []
    for _ in range(2):  # target 0 under the index bits 1 and 2: 8 gates -> a 2-bit multiplexor
        for c in (2, 1, 2, 1):
            gates += [G("Ry", (0,), (), float(rng.normal()), 1.0), G("X", (0,), (c,), None, 1.0)]
    for c in (2, 2):  # target 1 under index bit 2: too short for a table, stays plain
        gates += [G("Ry", (1,), (), float(rng.normal()), 1.0), G("X", (1,), (c,), None, 1.0)]
    gates += [G("Ry", (2,), (), float(rng.normal()), 1.0)]
    psi0 = random_state(n, rng)
    want = np_oracle.simulate_gates(gates, n, psi0)

    def passes():
        low = np.ascontiguousarray(lower_gates(gates, n), dtype=cabi.GATE_DTYPE)
        opts = cabi.PlanOpts.make(jit="off")
        lib = ctypes.CDLL(_build())
        out = np.zeros(256, dtype=np.int32)
        m = lib.emul_describe_segment(n, cabi.B2SV_C128, low.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(low)), ctypes.byref(opts), 0,
                                      out.ctypes.data_as(ctypes.c_void_p), len(out))
        return [int(out[2 * i]) for i in range(m)]

    got, _info = emul(gates, n, psi0)
    assert np.abs(got - want).max() < 1e-12
    kinds = passes()
    assert 5 not in kinds and kinds.count(6) == 1, kinds  # no MUX pass, one BLOCK pass (op codes of kernels.cuh)

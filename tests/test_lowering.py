"""Host logic: gate lowering and the Subspace program (CPU only)."""

import numpy as np
import pytest
import tequila as tq

from oracle import np_oracle
from tests.helpers import golden_names, load_golden, random_circuit
from unitaria_b200 import cabi
from unitaria_b200.lowering import LoweringError, lower_gates, touches_any_qubit
from unitaria_b200.subspace_prog import SubspaceProgram, to_spec


def lowered_matrix(rec):
    """2x2 unitary a lowered single-target gate stands for."""
    m = rec["m"]
    c = lambda i: complex(m[2 * i], m[2 * i + 1])  # noqa: E731
    op = int(rec["op"])
    if op == cabi.OP_MAT1:
        return np.array([[c(0), c(1)], [c(2), c(3)]])
    if op == cabi.OP_DIAG1:
        return np.array([[c(0), 0], [0, c(1)]])
    if op == cabi.OP_X:
        return np.array([[0, 1], [1, 0]], dtype=complex)
    raise AssertionError(op)


@pytest.mark.parametrize("seed", range(4))
def test_lowered_gates_match_oracle_unitaries(seed):
    rng = np.random.default_rng(seed)
    circuit = random_circuit(6, 300, rng)
    gates = [g for g in circuit.gates if g.name != "I"]
    low = lower_gates(gates, 6)
    assert len(low) == len(gates)
    for g, rec in zip(gates, low):
        mask = sum(1 << c for c in g.control)
        assert int(rec["ctrl_mask"]) == mask
        if g.name == "GlobalPhase":
            assert rec["op"] == cabi.OP_PHASE
            assert abs(complex(rec["m"][0], rec["m"][1]) - np.exp(1j * g.parameter)) < 1e-15
        elif g.name == "SWAP":
            assert rec["op"] == cabi.OP_SWAP and {int(rec["t0"]), int(rec["t1"])} == set(g.target)
        else:
            assert int(rec["t0"]) == g.target[0]
            want = np_oracle.gate_unitary(g.name, getattr(g, "parameter", None), getattr(g, "power", 1))
            assert np.abs(lowered_matrix(rec) - want).max() < 1e-15, g


def test_x_is_lowered_to_exact_data_movement_and_noops_are_exact():
    c = tq.gates.X(0) + tq.gates.CNOT(1, 0) + tq.gates.Toffoli(0, 1, 2) + tq.gates.I(3)
    low = lower_gates(c.gates, 4)
    assert list(low["op"]) == [cabi.OP_X] * 3
    zero = tq.gates.Ry(0.0, 1) + tq.gates.Rz(0.0, 1) + tq.gates.Phase(1, angle=0.0) + tq.gates.GlobalPhase(0.0) + tq.gates.Z(0, power=2)
    for rec in lower_gates(zero.gates, 2):
        m = rec["m"]
        if rec["op"] == cabi.OP_MAT1:
            assert list(m) == [1, 0, 0, 0, 0, 0, 1, 0]
        elif rec["op"] == cabi.OP_DIAG1:
            assert list(m[:4]) == [1, 0, 1, 0]
        else:
            assert list(m[:2]) == [1, 0]
    z = lower_gates(tq.gates.Z(0).gates, 1)[0]
    assert list(z["m"][:4]) == [1, 0, -1, 0]


def test_lowering_rejects_bad_gates():
    class G:
        name, target, control, parameter, power = "CCZZ", (0,), (), 0.0, 1

    with pytest.raises(LoweringError):
        lower_gates([G()], 2)
    with pytest.raises(LoweringError):
        lower_gates(tq.gates.X(5).gates, 3)


def test_touches_any_qubit():
    assert not touches_any_qubit(tq.gates.GlobalPhase(0.3).gates)
    assert touches_any_qubit(tq.gates.GlobalPhase(0.3).add_controls([1]).gates)
    assert touches_any_qubit(tq.gates.I(0).gates)


def test_golden_fixture_roundtrip_through_lowering():
    for name in golden_names(max_qubits=20)[:12]:
        case = load_golden(name)
        low = lower_gates(case.gates, case.n_qubits)
        assert len(low) == sum(1 for g in case.gates if g.name != "I")


# -- Subspace program: a Python mirror of the device walk (kernels.cuh sub_member_rank) ------------


def walk(prog: SubspaceProgram, bits: int):
    ins = prog.instrs
    pc, rank = prog.entry, 0
    for _ in range(len(ins) + 1):
        i = ins[pc]
        op, pos, ln = int(i["op"]), int(i["pos"]), int(i["len"])
        field = (bits >> pos) & ((1 << ln) - 1)
        if op == cabi.SUB_END:
            return rank
        if op == cabi.SUB_ZEROS:
            if field:
                return None
            pc = int(i["next0"])
        elif op == cabi.SUB_FREE:
            rank += field * int(i["weight"])
            pc = int(i["next0"])
        else:
            if (bits >> pos) & 1:
                rank += int(i["add"])
                pc = int(i["next1"])
            else:
                pc = int(i["next0"])
    raise AssertionError("program did not terminate")


@pytest.mark.parametrize("name", ["laplace_n3", "laplace_readme_n8", "bpx_L3", "bpx_L2", "pinv_bpx_L2", "gaussian_s3_t4", "cfg5_bits5"])
@pytest.mark.parametrize("which", ["subspace_in", "subspace_out"])
def test_subspace_program_reproduces_enumerate_basis_order(name, which):
    case = load_golden(name)
    spec = case.meta[which]
    prog = SubspaceProgram(spec)
    want = np_oracle.enumerate_basis(spec)
    assert prog.dimension == len(want) and prog.total_qubits == np_oracle.spec_total_qubits(spec)
    got = {}
    for i in range(2**prog.total_qubits):
        r = walk(prog, i)
        if r is not None:
            got[r] = i
    assert sorted(got) == list(range(len(want)))
    assert [got[r] for r in range(len(want))] == list(want)


@pytest.mark.parametrize("name", ["laplace_n3", "laplace_readme_n8", "bpx_L3", "bpx_L2", "pinv_bpx_L2", "gaussian_s3_t4", "cfg5_bits5"])
@pytest.mark.parametrize("which", ["subspace_in", "subspace_out"])
def test_spec_member_is_enumerate_basis_by_rank(name, which):
    from unitaria_b200.subspace_prog import spec_member

    spec = to_spec(load_golden(name).meta[which])
    want = np_oracle.enumerate_basis(spec)
    assert [spec_member(spec, r) for r in range(len(want))] == list(want)
    with pytest.raises(IndexError):
        spec_member(spec, len(want))


def test_index_window_subspaces_are_recognised():
    for text, want in [("00###", True), ("###", True), ("000", True), ("0#0", False), ("#0", False), ("##0", False)]:
        prog = SubspaceProgram(text)
        assert prog.is_index_window is want, text
        if want:
            assert list(np_oracle.enumerate_basis(to_spec(text))) == list(range(prog.dimension))
    assert SubspaceProgram(load_golden("laplace_readme_n8").meta["subspace_out"]).is_index_window is False


def test_subspace_string_form():
    # subspace.py:114-125: leftmost character is the most significant qubit
    assert to_spec("0#") == [[[], []], 0]
    p = SubspaceProgram("00###")
    assert p.total_qubits == 5 and p.dimension == 8
    assert len(p.instrs) == 3  # END, ZEROS(3,2), FREE(0,3): runs are merged
    assert [walk(p, i) for i in range(9)] == [0, 1, 2, 3, 4, 5, 6, 7, None]


def test_memoised_lowering_equals_gate_by_gate_lowering():
    """lower_gates lowers each distinct gate record once and gathers (sub-circuit reuse, qsvt.py:341-359): same
    bytes as lowering every gate on its own, I gates dropped, order kept, unhashable fields tolerated."""
    from tests.helpers import load_golden
    from unitaria_b200.fixtures import GateRecord
    from unitaria_b200.lowering import _lower_each

    case = load_golden("pinv_bpx_L2")
    gates = list(case.gates)
    gates.insert(7, GateRecord("I", (3,), (), None, 1.0))
    gates.insert(400, GateRecord("I", (5,), (), None, 1.0))
    gates.append(GateRecord("X", [2], [0, 1], None, 1.0))  # list-valued targets: not hashable
    low = lower_gates(gates, case.n_qubits)
    each = _lower_each(gates, case.n_qubits)
    assert len(low) == len(gates) - 2 and low.tobytes() == each.tobytes()

"""Host logic of the distributed layer on CPU: world_size-2 and -4 ``gloo`` runs.

The shard backend here is a TEST DOUBLE built on the CPU oracle (tests may use the oracle; the
product backend is the CUDA engine, exercised on GPUs by tests/test_gpu_distributed.py).  What is
under test is everything above it: layout tracking, gate localisation, rank relabels, swap
planning, the pack/exchange/unpack order and the global reduction.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import np_oracle
from tests.helpers import load_golden, random_circuit
from unitaria_b200 import cabi
from unitaria_b200.distributed import Layout, ShardedStateVector, TorchComm, localise_gate
from unitaria_b200.fixtures import GateRecord
from unitaria_b200.subspace_prog import SubspaceProgram, to_spec


class NumpyShard:
    supports_gather = False  # round-1 schedule: pack / exchange / unpack swaps

    def __init__(self, n_local):
        self.n = n_local
        self.psi = np.zeros(2**n_local, dtype=np.complex128)
        self.send = np.zeros(2 ** (n_local - 1), dtype=np.complex128)
        self.recv = np.zeros(2 ** (n_local - 1), dtype=np.complex128)

    def set_basis(self, index):
        self.psi[:] = 0
        self.psi[index] = 1

    def clear(self):
        self.psi[:] = 0

    def apply(self, gates, opts=None):
        for g in gates:
            np_oracle.apply_gate(self.psi, self.n, g)

    def _half(self, q, keep_bit):
        idx = np.arange(2**self.n)
        return idx[((idx >> q) & 1) != keep_bit]

    chunks = 2  # exercise the pipelined pack / exchange / unpack path of ShardedStateVector._swap

    @property
    def half(self):
        return self.psi.size // 2

    def pack(self, q, keep_bit, first=0, count=None):
        count = self.half - first if count is None else count
        self.send[first : first + count] = self.psi[self._half(q, keep_bit)[first : first + count]]
        s = torch.from_numpy(self.send.view(np.float64))[2 * first : 2 * (first + count)]
        r = torch.from_numpy(self.recv.view(np.float64))[2 * first : 2 * (first + count)]
        return s, r

    def unpack(self, q, keep_bit, first=0, count=None):
        count = self.half - first if count is None else count
        self.psi[self._half(q, keep_bit)[first : first + count]] = self.recv[first : first + count]

    def norm2_shard(self, sp):
        """Python mirror of the device walk over a per-rank specialised membership program."""
        ins = sp.instrs
        total = 0.0
        for i in range(2**self.n):
            pc = sp.entry
            ok = None
            for _ in range(len(ins) + 1):
                op, pos, ln = int(ins["op"][pc]), int(ins["pos"][pc]), int(ins["len"][pc])
                field = (i >> pos) & ((1 << ln) - 1)
                if op == cabi.SUB_END:
                    ok = True
                    break
                if op == cabi.SUB_REJECT:
                    ok = False
                    break
                if op == cabi.SUB_ZEROS:
                    if field:
                        ok = False
                        break
                    pc = int(ins["next0"][pc])
                elif op == cabi.SUB_FREE:
                    pc = int(ins["next0"][pc])
                else:
                    pc = int(ins["next1"][pc]) if (i >> pos) & 1 else int(ins["next0"][pc])
            assert ok is not None
            if ok:
                total += abs(self.psi[i]) ** 2
        return total

    def download(self):
        return self.psi.copy()

    def close(self):
        pass


class NumpyGatherShard(NumpyShard):
    """Test double of the fused gather exchange (b2sv_dist_perm): every rank sees all shards (all-gather over
    gloo stands in for the peer-mapped HBM) and evaluates out[j] = in[f^-1(j)] with the flip-flag semantics of
    include/b2sv.h -- which is what pins the host logic: run collection, relabels, k-qubit swaps."""

    supports_gather = True

    def __init__(self, n_local):
        super().__init__(n_local)
        self.sweeps = 0

    def prepare_dist(self, gates_pos, n_total):
        return list(gates_pos)

    def dist_perm(self, gates_pos, n_total, flip_in, flip_out, comm, opts=None):
        nl = self.n
        full = np.concatenate(comm.all_gather_array(self.psi))   # index = physical rank : local index
        j_logical = np.arange(2**n_total, dtype=np.int64) ^ (int(flip_out) << nl)
        src = j_logical.copy()
        for g in reversed(gates_pos):  # X and SWAP are self-inverse: f^-1 applies them in reverse order
            assert g.name.upper() in ("X", "SWAP"), g
            cm = 0
            for c in g.control:
                cm |= 1 << int(c)
            on = (src & cm) == cm
            if g.name.upper() == "X":
                src = np.where(on, src ^ (1 << int(g.target[0])), src)
            else:
                a, b = (int(t) for t in g.target)
                differ = ((src >> a) ^ (src >> b)) & 1 == 1
                src = np.where(on & differ, src ^ ((1 << a) | (1 << b)), src)
        out = full[src ^ (int(flip_in) << nl)]
        self.psi = out[comm.rank << nl : (comm.rank + 1) << nl].copy()
        self.sweeps += 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case_name, seed, out_dir, gather=False, choose=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if case_name.startswith("random"):
            n = int(case_name.split("_")[1])
            rng = np.random.default_rng(seed)
            gates = random_circuit(n, 160, rng, max_controls=3).gates
            spec = to_spec("0" + "#" * (n - 1))
            basis = int(rng.integers(0, 2 ** (n - 1)))
        else:
            case = load_golden(case_name)
            n, gates, spec = case.n_qubits, case.gates, case.meta["subspace_out"]
            basis = 3
        gq = None
        if choose:
            from unitaria_b200.distributed import choose_global_qubits

            gq = choose_global_qubits(gates, n, int(np.log2(world)), spec)
        sv = ShardedStateVector(n, TorchComm(), NumpyGatherShard if gather else NumpyShard, global_qubits=gq)
        assert sv.gather == gather
        sv.set_basis(basis)
        sv.apply(gates)
        norm2 = sv.project_norm2(SubspaceProgram(spec))
        swaps_before_restore = sv.stats["swaps"]  # the projection itself must not move data
        local = sv.local_amplitudes()
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), local)
        if rank == 0:
            np.save(os.path.join(out_dir, "meta.npy"), np.array([norm2, swaps_before_restore, sv.stats["swaps"], sv.stats["dist_sweeps"]]))
    finally:
        dist.destroy_process_group()


def run_case(tmp_path, world, case_name, seed=0, gather=False, choose=False):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case_name, seed, str(tmp_path), gather, choose), nprocs=world, join=True)
    full = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    meta = np.load(tmp_path / "meta.npy")
    return full, meta


@pytest.mark.parametrize("gather", [False, True])
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case_name", ["laplace_n4", "cfg5_bits5", "bpx_L2"])
def test_sharded_matches_single_register_goldens(tmp_path, world, case_name, gather):
    case = load_golden(case_name)
    full, meta = run_case(tmp_path, world, case_name, gather=gather)
    want = np_oracle.simulate_gates(case.gates, case.n_qubits, 3)
    assert np.abs(full - want).max() < 1e-12
    basis = np_oracle.enumerate_basis(case.meta["subspace_out"])
    assert abs(meta[0] - np.sum(np.abs(want[basis]) ** 2)) < 1e-12


@pytest.mark.parametrize("gather", [False, True])
@pytest.mark.parametrize("world,n,seed", [(2, 6, 1), (4, 7, 2), (2, 9, 3), (8, 8, 4)])
def test_sharded_matches_single_register_random(tmp_path, world, n, seed, gather):
    """Every gate kind, controls on global qubits, rank relabels (uncontrolled X on a global qubit),
    diagonal gates on global targets, global<->local swaps / distributed permutation sweeps."""
    full, meta = run_case(tmp_path, world, f"random_{n}", seed, gather=gather)
    rng = np.random.default_rng(seed)
    gates = random_circuit(n, 160, rng, max_controls=3).gates
    basis = int(rng.integers(0, 2 ** (n - 1)))
    want = np_oracle.simulate_gates(gates, n, basis)
    assert np.abs(full - want).max() < 1e-12
    member = np_oracle.enumerate_basis(to_spec("0" + "#" * (n - 1)))
    assert abs(meta[0] - np.sum(np.abs(want[member]) ** 2)) < 1e-12
    assert meta[1] > 0  # the circuit did need global<->local swaps
    if gather:
        assert meta[3] > 0  # ... and they ran as distributed sweeps


def test_cfg5_needs_no_swap_with_the_gather_exchange(tmp_path):
    """BASELINE config 5's shape (Increment across the rank boundary & low-qubit state preparation): the
    Increment's X-family gates on global targets ride in distributed sweeps, no qubit is ever swapped."""
    case = load_golden("cfg5_bits5")
    full, meta = run_case(tmp_path, 4, "cfg5_bits5", gather=True)
    want = np_oracle.simulate_gates(case.gates, case.n_qubits, 3)
    assert np.abs(full - want).max() < 1e-12
    assert meta[1] == 0 and 1 <= meta[3] <= 3, meta


@pytest.mark.parametrize("case_name,world", [("cfg5_bits5", 4), ("bpx_L2", 2), ("random_8", 4)])
def test_chosen_global_qubits(tmp_path, case_name, world):
    """Initial layout picked by choose_global_qubits (no dense targets, no post-selected ancilla in the rank number):
    same amplitudes, same norm."""
    full, meta = run_case(tmp_path, world, case_name, seed=5, gather=True, choose=True)
    if case_name.startswith("random"):
        rng = np.random.default_rng(5)
        n = 8
        gates = random_circuit(n, 160, rng, max_controls=3).gates
        basis = int(rng.integers(0, 2 ** (n - 1)))
        want = np_oracle.simulate_gates(gates, n, basis)
        members = np_oracle.enumerate_basis(to_spec("0" + "#" * (n - 1)))
    else:
        case = load_golden(case_name)
        want = np_oracle.simulate_gates(case.gates, case.n_qubits, 3)
        members = np_oracle.enumerate_basis(case.meta["subspace_out"])
    assert np.abs(full - want).max() < 1e-12
    assert abs(meta[0] - np.sum(np.abs(want[members]) ** 2)) < 1e-12


def test_choose_global_qubits_rules():
    from unitaria_b200.distributed import choose_global_qubits

    case = load_golden("cfg5_bits5")  # qubits 0-5 state preparation (dense), 6-10 Increment, 11 borrowed ancilla
    assert choose_global_qubits(case.gates, 12, 2, case.meta["subspace_out"]) == [9, 10]
    assert choose_global_qubits(case.gates, 12, 2, None) == [10, 11]


def test_localise_gate_rules():
    lay = Layout(6, 2)  # qubits 4,5 global
    # control on a global qubit: dropped on ranks where it is 1, gate skipped where it is 0
    g = GateRecord("X", (0,), (1, 5), 1.0, 1.0)
    assert localise_gate(g, lay, rank=0) is None
    lg = localise_gate(g, lay, rank=2)
    assert lg.target == (0,) and lg.control == (1,)
    # diagonal gate on a global target: per-rank scalar on the remaining controls
    rz = GateRecord("Rz", (4,), (2,), 0.6, 1.0)
    a, b = localise_gate(rz, lay, 0), localise_gate(rz, lay, 1)
    assert a.name == "GlobalPhase" and a.control == (2,) and abs(a.parameter + 0.3) < 1e-15 and abs(b.parameter - 0.3) < 1e-15
    ph = GateRecord("Phase", (5,), (), 0.4, 1.0)
    assert localise_gate(ph, lay, 0) is None and localise_gate(ph, lay, 2).parameter == 0.4
    # rank relabel: flip flag inverts the meaning of the rank bit
    lay.flip[lay.pos[5]] = 1
    assert localise_gate(g, lay, rank=0) is not None and localise_gate(g, lay, rank=2) is None

"""The sharded path on real GPUs.

* ``test_two_ranks_on_one_gpu_*`` run everywhere a single B200 is present: two processes (gloo control plane)
  share device 0, map each other's shards with CUDA IPC and run the fused gather exchange
  (``b2sv_dist_perm``), the per-rank projection programs (``b2sv_project_norm2_at``) and the round-1 pack /
  unpack kernels -- no kernel ever waits for another rank (ordering is by interprocess events), so time-slicing
  two ranks on one device is safe.
* ``test_sharded_gpu_matches_oracle`` needs >= 2 B200s (``gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_distributed.py``):
  one process per GPU, NCCL default group.

Results are checked against the CPU oracle on the full register."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle, np_oracle
from tests.helpers import load_golden, random_circuit
from unitaria_b200 import cabi
from unitaria_b200.cabi import PlanOpts
from unitaria_b200.distributed import ShardedStateVector, TorchComm, make_gpu_shard_factory
from unitaria_b200.subspace_prog import SubspaceProgram, to_spec

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case(case_name, seed):
    if case_name.startswith("random"):
        n = int(case_name.split("_")[1])
        rng = np.random.default_rng(seed)
        gates = random_circuit(n, 240, rng, max_controls=3).gates
        return n, gates, to_spec("0" + "#" * (n - 1))
    case = load_golden(case_name)
    return case.n_qubits, case.gates, case.meta["subspace_out"]


def _worker(rank, world, port, case_name, dtype, basis, out_dir, exchange, backend, one_device, jit, seed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    device = 0 if one_device else rank
    torch.cuda.set_device(device)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{device}"))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, gates, spec = _case(case_name, seed)
        comm = TorchComm(device=torch.device(f"cuda:{device}"))
        sv = ShardedStateVector(n, comm, make_gpu_shard_factory(dtype, device, exchange))
        assert sv.gather == (exchange == "gather"), getattr(sv.shard, "attach_error", "")
        opts = PlanOpts.make(jit=jit)
        norm2 = 0.0
        for b in (basis + 1, basis):  # twice: the second pass replays the compiled program on a dirty register
            sv.set_basis(b)
            sv.apply(gates, opts)
            norm2 = sv.project_norm2(SubspaceProgram(spec))
        stats = dict(sv.stats)
        launches = sv.shard.sv.stats()["launches_by_class"]
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), np.array(sv.local_amplitudes()))
        if rank == 0:
            np.save(os.path.join(out_dir, "meta.npy"), np.array([norm2, stats["swaps"], stats["dist_sweeps"], launches["dist_perm"], launches["swap"]]))
        sv.close()
    finally:
        dist.destroy_process_group()


def _run(tmp_path, world, case_name, dtype, exchange, backend, one_device, jit="sync", seed=0, basis=5):
    out = tmp_path / f"{case_name}_{world}_{exchange}_{jit}"
    out.mkdir()
    mp.spawn(_worker, args=(world, _free_port(), case_name, dtype, basis, str(out), exchange, backend, one_device, jit, seed), nprocs=world, join=True)
    full = np.concatenate([np.load(out / f"shard{r}.npy") for r in range(world)])
    return full, np.load(out / "meta.npy")


def _want(case_name, seed, basis=5):
    n, gates, spec = _case(case_name, seed)
    want = c_oracle.simulate_gates(gates, n, basis)
    members = np_oracle.enumerate_basis(spec)
    return want, float(np.sum(np.abs(want[members]) ** 2))


@pytest.mark.parametrize(
    "case_name,dtype,tol,world,jit",
    [
        ("cfg5_bits13", "complex128", 1e-12, 2, "sync"),   # Increment across the rank boundary: no swap at all
        ("cfg5_bits13", "complex128", 1e-12, 4, "off"),    # ... interpreter form of the gather kernel
        ("laplace_n16", "complex128", 1e-12, 2, "sync"),   # Ry on the top (global) LCU qubits: real k-qubit swaps
        ("laplace_n16", "complex128", 1e-12, 4, "sync"),
        ("bpx_L4", "complex64", 1e-5, 2, "sync"),
        ("random_15", "complex128", 1e-12, 2, "sync"),     # every gate kind, relabels, global diagonal targets
        ("random_15", "complex128", 1e-12, 4, "sync"),
        ("random_12", "complex128", 1e-12, 2, "sync"),     # 11 local qubits: shards too small for the bit-sliced kernel
    ],
)
def test_two_ranks_on_one_gpu_gather_exchange(tmp_path, case_name, dtype, tol, world, jit):
    if cabi.device_count() < 1:
        pytest.fail("needs a CUDA device")
    want, want_norm2 = _want(case_name, 7)
    full, meta = _run(tmp_path, world, case_name, dtype, "gather", "gloo", True, jit=jit, seed=7)
    assert np.abs(full - want).max() <= tol, (case_name, world)
    assert abs(meta[0] - want_norm2) <= tol * 10
    assert meta[3] >= 1, "the distributed gather kernel did not run"
    if case_name.startswith("cfg5"):
        assert meta[1] == 0, "config 5 must not need a qubit swap with the gather exchange"


def test_two_ranks_on_one_gpu_round1_pack_kernels(tmp_path):
    """The pack / unpack kernels of the NCCL exchange (kept as the baseline transport), driven over gloo."""
    if cabi.device_count() < 1:
        pytest.fail("needs a CUDA device")
    want, want_norm2 = _want("laplace_n16", 7)
    full, meta = _run(tmp_path, 2, "laplace_n16", "complex128", "nccl", "gloo", True, seed=7)
    assert np.abs(full - want).max() <= 1e-12
    assert abs(meta[0] - want_norm2) <= 1e-11
    assert meta[1] >= 1 and meta[4] >= 2 and meta[3] == 0


def _worlds():
    n = cabi.device_count()
    worlds = [w for w in (2, 4, 8) if w <= n]
    only = os.environ.get("B2SV_TEST_WORLDS")  # e.g. "8": keep an 8-GPU lease short
    if only:
        worlds = [w for w in worlds if str(w) in only.split(",")]
    return worlds


@pytest.mark.parametrize("case_name,dtype,tol", [("cfg5_bits13", "complex128", 1e-12), ("laplace_n16", "complex128", 1e-12), ("bpx_L4", "complex64", 1e-5)])
def test_sharded_gpu_matches_oracle(tmp_path, case_name, dtype, tol):
    worlds = _worlds()
    if not worlds:
        pytest.skip("needs at least 2 GPUs")
    want, want_norm2 = _want(case_name, 0)
    for world, exchange in [(w, e) for w in worlds for e in ("gather", "nccl")]:
        full, meta = _run(tmp_path, world, case_name, dtype, exchange, "nccl", False)
        assert np.abs(full - want).max() <= tol, (case_name, world, exchange)
        assert abs(meta[0] - want_norm2) <= tol * 10, (case_name, world, exchange)

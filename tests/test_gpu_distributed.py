"""The sharded path on real GPUs (needs >= 2 B200s: run with ``gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_distributed.py``).

One process per GPU, NCCL; results are checked against the CPU oracle on the full register, and
against the single-GPU engine (shard invariance)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import c_oracle, np_oracle
from tests.helpers import load_golden
from unitaria_b200 import cabi
from unitaria_b200.distributed import ShardedStateVector, TorchComm, make_gpu_shard_factory
from unitaria_b200.subspace_prog import SubspaceProgram

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case_name, dtype, basis, out_dir, peer):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        case = load_golden(case_name)
        comm = TorchComm(device=torch.device(f"cuda:{rank}"))
        sv = ShardedStateVector(case.n_qubits, comm, make_gpu_shard_factory(dtype, rank))
        if peer:
            assert comm.enable_peer_pull(sv.shard.send), getattr(comm, "peer_error", "peer mapping failed")
        sv.set_basis(basis)
        sv.apply(case.gates)
        swaps = sv.stats["swaps"]
        norm2 = sv.project_norm2(SubspaceProgram(case.meta["subspace_out"]))
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), np.array(sv.local_amplitudes()))
        if rank == 0:
            np.save(os.path.join(out_dir, "meta.npy"), np.array([norm2, swaps]))
        sv.close()
    finally:
        dist.destroy_process_group()


def _worlds():
    n = cabi.device_count()
    return [w for w in (2, 4, 8) if w <= n]


@pytest.mark.parametrize("case_name,dtype,tol", [("cfg5_bits13", "complex128", 1e-12), ("laplace_n16", "complex128", 1e-12), ("bpx_L4", "complex64", 1e-5)])
def test_sharded_gpu_matches_oracle(tmp_path, case_name, dtype, tol):
    worlds = _worlds()
    if not worlds:
        pytest.skip("needs at least 2 GPUs")
    case = load_golden(case_name)
    basis = 5
    want = c_oracle.simulate_gates(case.gates, case.n_qubits, basis)
    members = np_oracle.enumerate_basis(case.meta["subspace_out"])
    want_norm2 = float(np.sum(np.abs(want[members]) ** 2))
    for world, peer in [(w, p) for w in worlds for p in (False, True)]:
        out = tmp_path / f"w{world}_{int(peer)}"
        out.mkdir()
        mp.spawn(_worker, args=(world, _free_port(), case_name, dtype, basis, str(out), peer), nprocs=world, join=True)
        full = np.concatenate([np.load(out / f"shard{r}.npy") for r in range(world)])
        meta = np.load(out / "meta.npy")
        assert np.abs(full - want).max() <= tol, (case_name, world)
        assert abs(meta[0] - want_norm2) <= tol * 10, (case_name, world)

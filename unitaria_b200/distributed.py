"""Distributed layer: one statevector sharded over 2/4/8 B200s of one box by *global qubits*.

The reference is single-process (no distributed code at all), so this layer is
new.  It keeps the seam's semantics: ``ShardedStateVector.apply(gates)`` on G ranks gives the same
amplitudes as ``Circuit.simulate`` (/root/reference/src/unitaria/circuit.py:42-70) on one register.

Layout.  The 2^n amplitudes are split by the top ``g = log2(G)`` *positions* of the index: rank r owns
the contiguous shard whose global position bits equal r (one process per GPU, ``n - g`` local
qubits).  A logical qubit may sit at any position; ``Layout`` tracks the permutation.

Per gate:
* non-diagonal targets must be local.  Controls on global positions are free: the rank either
  satisfies them (the control is dropped) or skips the gate.  Diagonal gates on a global target
  become a per-rank scalar phase on the remaining controls.  An uncontrolled X on a global
  position is a rank relabel (a flip flag), no data moves.
* when a gate needs a global qubit as a target, the gates gathered so far are flushed to the local
  engine (which fuses them as usual) and the qubit is swapped with a local one that is not needed
  soon: every rank packs the half of its shard whose local bit differs from its rank bit
  (``b2sv_swap_pack``), exchanges it with rank ``r ^ bit`` over NCCL (grouped send/recv -- the
  pairwise case of an all-to-all) and unpacks in place.  Half a shard crosses NVLink per swap.

The local engine is reached through a small backend interface so that the host logic (layout,
gate localisation, swap planning, exchange order) is testable on CPU with ``gloo`` and a test
double; the product backend is :class:`GpuShard` (the CUDA engine, no fallback).
"""

from __future__ import annotations

import math
import os
import time

import numpy as np

from .fixtures import GateRecord

_DIAGONAL = {"Z", "RZ", "PHASE", "GLOBALPHASE", "I"}


class Layout:
    """logical qubit -> position.  Positions < n_local are bits of the local index, the rest are rank bits."""

    def __init__(self, n_qubits: int, n_global: int):
        self.n = n_qubits
        self.g = n_global
        self.n_local = n_qubits - n_global
        self.pos = list(range(n_qubits))   # pos[q]
        self.at = list(range(n_qubits))    # at[p] = q
        self.flip = [0] * n_qubits          # flip flag per GLOBAL position (rank relabels)

    def is_local(self, q: int) -> bool:
        return self.pos[q] < self.n_local

    def swap_positions(self, pa: int, pb: int) -> None:
        qa, qb = self.at[pa], self.at[pb]
        self.at[pa], self.at[pb] = qb, qa
        self.pos[qa], self.pos[qb] = pb, pa

    def rank_bit(self, rank: int, q: int) -> int:
        """Logical value of global qubit q on this rank."""
        p = self.pos[q]
        return ((rank >> (p - self.n_local)) & 1) ^ self.flip[p]


def _is_diagonal(g) -> bool:
    name = g.name.upper()
    if name in _DIAGONAL:
        return True
    return False


def _nd_targets(g):
    """Qubits on which the gate acts non-diagonally (must be local)."""
    return () if _is_diagonal(g) else tuple(int(t) for t in g.target)


def localise_gate(g, layout: Layout, rank: int):
    """Translate one gate into the shard's coordinates for `rank`: a GateRecord on local bit
    positions, or None when this rank does not take part.  Requires the gate's non-diagonal
    targets to be local."""
    name = g.name.upper()
    target = tuple(int(t) for t in (g.target or ()))
    control = tuple(int(c) for c in (g.control or ()))
    local_ctrl = []
    for c in control:
        if layout.is_local(c):
            local_ctrl.append(layout.pos[c])
        elif layout.rank_bit(rank, c) == 0:
            return None  # a global control is 0 on this rank
    param = float(getattr(g, "parameter", 0.0) or 0.0)
    power = getattr(g, "power", 1)
    power = 1.0 if power is None else float(power)
    if name == "I":
        return None
    if name == "GLOBALPHASE":
        return GateRecord("GlobalPhase", (), tuple(local_ctrl), param, 1.0)
    if name in ("Z", "RZ", "PHASE") and not layout.is_local(target[0]):
        # diagonal gate on a global qubit: this rank sees one of the two diagonal entries
        bit = layout.rank_bit(rank, target[0])
        if name == "RZ":
            angle = (param / 2.0) if bit else (-param / 2.0)
        elif name == "PHASE":
            angle = param if bit else 0.0
        else:
            angle = math.pi * power if bit else 0.0
        if angle == 0.0:
            return None
        return GateRecord("GlobalPhase", (), tuple(local_ctrl), angle, 1.0)
    lt = tuple(layout.pos[t] for t in target)
    assert all(p < layout.n_local for p in lt), "non-diagonal target is not local"
    if name in ("RX", "RY", "RZ", "PHASE"):
        return GateRecord(g.name, lt, tuple(local_ctrl), param, 1.0)
    return GateRecord(g.name, lt, tuple(local_ctrl), power, power)


class TorchComm:
    """torch.distributed plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests)."""

    def __init__(self, device=None):
        import torch.distributed as dist

        self.dist = dist
        self.rank = dist.get_rank()
        self.size = dist.get_world_size()
        self.device = device

    def exchange(self, send, recv, partner: int) -> None:
        self.wait(self.start_exchange(send, recv, partner), send)

    def start_exchange(self, send, recv, partner: int):
        """Grouped send/recv with one peer (the pairwise case of an all-to-all); returns without waiting."""
        dist = self.dist
        return dist.batch_isend_irecv([dist.P2POp(dist.isend, send, partner), dist.P2POp(dist.irecv, recv, partner)])

    def wait(self, reqs, tensor) -> None:
        for r in reqs:
            r.wait()
        if tensor.is_cuda:
            import torch

            # r.wait() only makes torch's current stream wait for the transfer; block the host on that
            # stream (not the whole device: later pieces of a pipelined exchange are still in flight)
            torch.cuda.current_stream(tensor.device).synchronize()

    # -- peer-memory exchange: pull the partner's packed half straight out of its HBM over NVLink ------
    def enable_peer_pull(self, send_tensor) -> bool:
        """Map every rank's send buffer into this process (CUDA IPC through torch's storage sharing) so
        that an exchange is ONE device-to-device copy over NVLink issued by the receiver, with no NCCL
        channel kernels in between.  Returns False (and leaves the NCCL path active) if mapping fails."""
        import torch

        self.peer_send = {}
        try:
            handle = send_tensor.untyped_storage()._share_cuda_()
            gathered = [None] * self.size
            self.dist.all_gather_object(gathered, (handle, int(send_tensor.storage_offset()), int(send_tensor.numel())))
            for r, (h, off, n) in enumerate(gathered):
                if r == self.rank:
                    continue
                storage = torch.UntypedStorage._new_shared_cuda(*h)
                self.peer_send[r] = torch.empty(0, dtype=torch.uint8, device=storage.device).set_(storage, off, (n,))
            ok = 1.0
        except Exception as exc:  # pragma: no cover - depends on the platform
            self.peer_error = repr(exc)
            ok = 0.0
        all_ok = self.allreduce_min(ok) > 0.5
        if not all_ok:
            self.peer_send = {}
        return all_ok

    def pull_async(self, recv_view, partner: int, first_byte: int):
        """Enqueue recv_view <- partner's send buffer [first_byte, +len) and return an event that fires when
        the bytes have landed (torch makes the destination device's current stream wait for a cross-device
        copy, so an event recorded there afterwards covers it)."""
        import torch

        n = recv_view.numel()
        recv_view.copy_(self.peer_send[partner][first_byte : first_byte + n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(recv_view.device))
        return ev

    def allreduce_min(self, value: float) -> float:
        import torch

        t = torch.tensor([value], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def allreduce_sum(self, value: float) -> float:
        import torch

        t = torch.tensor([value], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t)
        return float(t.item())

    def allreduce_max(self, value: float) -> float:
        import torch

        t = torch.tensor([value], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self) -> None:
        self.dist.barrier()


class GpuShard:
    """Product backend: the shard lives in the CUDA engine; exchange buffers are torch tensors
    (tensor hand-off only: NCCL needs them), filled and drained by the engine's pack kernels."""

    def __init__(self, n_local: int, dtype: str, device: int):
        import torch

        from .engine import StateVector

        self.device = device
        self.sv = StateVector(n_local, dtype=dtype, device=device)
        half_bytes = (self.sv.size // 2) * self.sv.amp_bytes
        self.send = torch.empty(half_bytes, dtype=torch.uint8, device=f"cuda:{device}")
        self.recv = torch.empty(half_bytes, dtype=torch.uint8, device=f"cuda:{device}")
        torch.cuda.synchronize(device)

    def set_basis(self, index: int):
        self.sv.set_basis(index)

    def clear(self):
        self.sv.set_zero()

    def apply(self, gates, opts=None):
        if gates:
            self.sv.apply(gates, opts)

    chunks = 8  # pack / exchange / unpack are pipelined over this many pieces of the half shard
    min_chunk = 1 << 16

    def pack(self, q: int, keep_bit: int, first: int = 0, count: int | None = None):
        """Pack amplitudes [first, first+count) of the moving half; returns the matching views of the
        send / receive buffers (bytes)."""
        half = self.sv.size >> 1
        count = half - first if count is None else count
        self.sv.swap_pack(q, keep_bit, self.send.data_ptr(), first, count)
        b = self.sv.amp_bytes
        return self.send[first * b : (first + count) * b], self.recv[first * b : (first + count) * b]

    def unpack(self, q: int, keep_bit: int, first: int = 0, count: int | None = None):
        self.sv.swap_unpack(q, keep_bit, self.recv.data_ptr(), first, count)

    @property
    def half(self) -> int:
        return self.sv.size >> 1

    @property
    def amp_bytes(self) -> int:
        return self.sv.amp_bytes

    def views(self, first: int, count: int):
        b = self.sv.amp_bytes
        return self.send[first * b : (first + count) * b], self.recv[first * b : (first + count) * b]

    def norm2_shard(self, shard_prog) -> float:
        return self.sv.project_norm2_shard(shard_prog)

    def download(self) -> np.ndarray:
        return self.sv.download()

    def close(self):
        self.sv.close()


class ShardedStateVector:
    """A 2^n_qubits-amplitude register over ``comm.size`` ranks (power of two)."""

    def __init__(self, n_qubits: int, comm, shard_factory, lookahead: int = 4096):
        g = int(math.log2(comm.size))
        if 1 << g != comm.size:
            raise ValueError("world size must be a power of two")
        if n_qubits - g < 2:
            raise ValueError("too few local qubits")
        self.n = n_qubits
        self.comm = comm
        self.rank = comm.rank
        self.layout = Layout(n_qubits, g)
        self.shard = shard_factory(n_qubits - g)
        self.lookahead = lookahead
        self.stats = {"swaps": 0, "swap_bytes": 0, "swap_s": 0.0, "flushes": 0, "local_gates": 0}

    # -- state ----------------------------------------------------------------------------------
    def set_basis(self, index: int) -> None:
        """|index> of the LOGICAL register (identity layout is restored first)."""
        self.layout = Layout(self.n, self.layout.g)
        nl = self.layout.n_local
        owner, local = index >> nl, index & ((1 << nl) - 1)
        if owner == self.rank:
            self.shard.set_basis(local)
        else:
            self.shard.clear()

    # -- gates ----------------------------------------------------------------------------------
    def apply(self, gates, opts=None) -> None:
        gates = [g for g in gates if g.name.upper() != "I"]
        lay = self.layout
        pending: list = []
        # next index at which each qubit is needed as a non-diagonal target (Belady eviction)
        need = _next_use_table(gates, self.n)

        def flush():
            if pending:
                self.shard.apply(pending, opts)
                self.stats["flushes"] += 1
                self.stats["local_gates"] += len(pending)
                pending.clear()

        for i, g in enumerate(gates):
            name = g.name.upper()
            nd = _nd_targets(g)
            # uncontrolled X on a global qubit: rank relabel
            if name == "X" and not g.control and getattr(g, "power", 1) in (1, None) and not lay.is_local(nd[0]):
                lay.flip[lay.pos[nd[0]]] ^= 1
                continue
            missing = [q for q in nd if not lay.is_local(q)]
            if missing:
                flush()
                for q in missing:
                    victim = self._pick_victim(need, i, exclude=set(nd))
                    self._swap(q, victim)
                # While the engine is drained anyway, also bring in every other global qubit that will be
                # a target later, as long as a local qubit with no further use can take its place: one
                # flush instead of one per qubit keeps long fused runs (permutation sweeps) in one piece.
                never = 1 << 60
                for p in range(lay.n_local, lay.n):
                    q = lay.at[p]
                    if need.next_after(q, i) >= never:
                        continue
                    victim = self._pick_victim(need, i, exclude=set(nd))
                    if victim is None or need.next_after(victim, i) < never:
                        break
                    self._swap(q, victim)
            lg = localise_gate(g, lay, self.rank)
            if lg is not None:
                pending.append(lg)
        flush()

    def _pick_victim(self, need, i: int, exclude) -> int:
        """Local qubit whose next use as a non-diagonal target is farthest away."""
        lay = self.layout
        best, best_next = None, -1
        for p in range(lay.n_local - 1, -1, -1):  # prefer high local positions on ties (cheap, long strides)
            q = lay.at[p]
            if q in exclude:
                continue
            nxt = need.next_after(q, i)
            if nxt > best_next:
                best, best_next = q, nxt
        return best

    def _swap(self, q_global: int, q_local: int) -> None:
        """Exchange the positions of a global and a local qubit (half a shard crosses the link)."""
        lay = self.layout
        pg, pl = lay.pos[q_global], lay.pos[q_local]
        bit = pg - lay.n_local
        my_bit = (self.rank >> bit) & 1  # PHYSICAL rank bit: data placement ignores the flip flag
        t0 = time.perf_counter()
        partner = self.rank ^ (1 << bit)
        chunks = getattr(self.shard, "chunks", 1)
        if getattr(self.comm, "peer_send", None):
            # peer-memory path: everybody packs, then every rank pulls its incoming half directly out of
            # the partner's HBM -- one device-to-device copy over NVLink, no NCCL channel kernels
            half = self.shard.half
            b = self.shard.amp_bytes
            ta = time.perf_counter()
            self.shard.pack(pl, my_bit)
            tb = time.perf_counter()
            self.comm.allreduce_sum(0.0)              # light barrier: the partner's send buffer is complete
            tc = time.perf_counter()
            # all pieces are queued on the copy engine at once; piece c is unpacked while c+1.. still fly
            step = max(1, half // max(1, chunks))
            pending = []
            for first in range(0, half, step):
                count = min(step, half - first)
                _s, recv = self.shard.views(first, count)
                pending.append((self.comm.pull_async(recv, partner, first * b), first, count))
            unpack_s = 0.0
            for ev, first, count in pending:
                ev.synchronize()
                tu = time.perf_counter()
                self.shard.unpack(pl, my_bit, first, count)
                unpack_s += time.perf_counter() - tu
            te = time.perf_counter()
            td = te - unpack_s  # transfer time not hidden behind unpacking
            self.comm.allreduce_sum(0.0)              # nobody re-packs while a peer is still reading
            tf = time.perf_counter()
            for key, dt in (("pack_s", tb - ta), ("barrier_s", tc - tb + tf - te), ("transfer_s", td - tc), ("unpack_s", te - td)):
                self.stats[key] = self.stats.get(key, 0.0) + dt
            send_bytes = half * b
        elif chunks > 1 and hasattr(self.comm, "start_exchange") and self.shard.half >= chunks * getattr(self.shard, "min_chunk", 1):
            # software pipeline: while NCCL moves piece c, the engine packs piece c+1 and unpacks c-1
            half = self.shard.half
            step = half // chunks
            sent = 0
            inflight = []
            for c in range(chunks):
                first = c * step
                count = step if c + 1 < chunks else half - first
                send, recv = self.shard.pack(pl, my_bit, first, count)
                inflight.append((self.comm.start_exchange(send, recv, partner), send, first, count))
                sent += int(send.numel() * send.element_size())
                if len(inflight) > 1:
                    reqs, tensor, f0, n0 = inflight.pop(0)
                    self.comm.wait(reqs, tensor)
                    self.shard.unpack(pl, my_bit, f0, n0)
            for reqs, tensor, f0, n0 in inflight:
                self.comm.wait(reqs, tensor)
                self.shard.unpack(pl, my_bit, f0, n0)
            send_bytes = sent
        else:
            send, recv = self.shard.pack(pl, my_bit)
            self.comm.exchange(send, recv, partner)
            self.shard.unpack(pl, my_bit)
            send_bytes = int(send.numel() * send.element_size())
        # the local position now carries the former global qubit with the global position's flip
        # flag folded into it: physical bit value = logical ^ flip  ->  restore with a local X
        if lay.flip[pg]:
            self.shard.apply([GateRecord("X", (pl,), (), 1.0, 1.0)])
            lay.flip[pg] = 0
        lay.swap_positions(pg, pl)
        self.stats["swaps"] += 1
        self.stats["swap_bytes"] += send_bytes
        self.stats["swap_s"] += time.perf_counter() - t0

    def restore_identity_layout(self) -> None:
        """Bring every logical qubit back to position == qubit (needed before projections / download)."""
        lay = self.layout
        # clear flip flags: an X on the data, done by making the qubit local once
        for p in range(lay.n_local, lay.n):
            if lay.flip[p]:
                q = lay.at[p]
                victim = lay.at[lay.n_local - 1]
                self._swap(q, victim)      # folds the flip into a local X
                self._swap(victim, q)      # and back
        # global positions first, then fix the local permutation with local SWAP gates
        for p in range(lay.n_local, lay.n):
            q_want = p
            if lay.at[p] != q_want:
                if lay.is_local(q_want):
                    self._swap(lay.at[p], q_want)
                else:  # wanted qubit sits at another global position: route through a local slot
                    tmp = lay.at[lay.n_local - 1]
                    self._swap(q_want, tmp)
                    self._swap(lay.at[p], q_want)
        swaps = []
        for p in range(lay.n_local):
            while lay.at[p] != p:
                other = lay.pos[p]
                swaps.append(GateRecord("SWAP", (p, other), (), 1.0, 1.0))
                lay.swap_positions(p, other)
        if swaps:
            self.shard.apply(swaps)

    # -- results ----------------------------------------------------------------------------------
    def project_norm2(self, subspace) -> float:
        """Global sum of |psi_i|^2 over the subspace with all higher qubits 0 (nodes/node.py:379,402).

        No data moves: every rank evaluates a membership program specialised to the current layout
        (qubits that sit in the rank number are resolved on the host) and the partial sums are
        all-reduced."""
        from .subspace_prog import ShardProgram, SubspaceProgram

        prog = subspace if isinstance(subspace, SubspaceProgram) else SubspaceProgram(subspace)
        lay = self.layout
        rank_value = lambda p: ((self.rank >> (p - lay.n_local)) & 1) ^ lay.flip[p]  # noqa: E731
        sp = ShardProgram(prog, lay.pos, lay.n_local, rank_value, self.n)
        local = 0.0 if sp.rejects_all else self.shard.norm2_shard(sp)
        return self.comm.allreduce_sum(local)

    def local_amplitudes(self) -> np.ndarray:
        """This rank's shard in identity layout: amplitudes (rank << n_local) | i."""
        self.restore_identity_layout()
        return self.shard.download()

    def close(self):
        if getattr(self.comm, "peer_send", None):
            self.comm.barrier()
            self.comm.peer_send = {}  # drop the IPC mappings before the owners free their buffers
            self.comm.barrier()
        self.shard.close()


class _NextUse:
    def __init__(self, uses):
        self.uses = uses  # per qubit: sorted gate indices where it is a non-diagonal target
        self.cursor = [0] * len(uses)

    def next_after(self, q: int, i: int) -> int:
        u = self.uses[q]
        c = self.cursor[q]
        while c < len(u) and u[c] <= i:
            c += 1
        self.cursor[q] = c
        return u[c] if c < len(u) else 1 << 60


def _next_use_table(gates, n: int) -> _NextUse:
    uses = [[] for _ in range(n)]
    for i, g in enumerate(gates):
        for q in _nd_targets(g):
            uses[q].append(i)
    return _NextUse(uses)


# ---------------------------------------------------------------------------------------------
# multi-GPU benchmark leg (bench.py --gpus N, launched by torchrun: one rank per GPU)
# ---------------------------------------------------------------------------------------------


def make_gpu_shard_factory(dtype: str, device: int):
    return lambda n_local: GpuShard(n_local, dtype, device)


def _family_single_gpu(case1, dtype, device, jit):
    """gates/s of the scaling family's single-GPU member (same amplitudes per GPU), 3 timed steps."""
    from .cabi import PlanOpts
    from .engine import StateVector
    from .lowering import lower_gates
    from .subspace_prog import SubspaceProgram

    n1 = max(1, case1.n_qubits)
    low = lower_gates(case1.gates, n1)
    prog = SubspaceProgram(case1.meta["subspace_out"])
    opts = PlanOpts.make(jit=jit)
    with StateVector(n1, dtype=dtype, device=device) as sv:
        for j in (1, 2, 3):
            sv.set_basis(j)
            sv.apply_lowered(low, opts)
            sv.project_norm2(prog)
        sv.sync()
        sv.mark(0)
        steps = 3
        for j in range(steps):
            sv.set_basis(5 + j)
            sv.apply_lowered(low, opts)
            sv.project_norm2(prog)
        sv.mark(1)
        ms = sv.elapsed_ms(0, 1)
        live = sv.stats()["gates_live"] // (steps + 3)
    return {"workload": case1.meta["name"], "qubits": n1, "gates_live": int(live), "ms_per_step": ms / steps, "value": live * steps / (ms * 1e-3), "unit": "gates/s"}


def bench_main(args, case, config_dict, cpu_sample, ClockSampler, measured_peak):
    import json

    import torch
    import torch.distributed as dist

    from .cabi import PlanOpts
    from .subspace_prog import SubspaceProgram

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    torch.cuda.set_device(local_rank)
    # the exchanges are big pairwise send/recv: give NCCL enough point-to-point channels to fill NVLink 5
    os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "32")
    os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    comm = TorchComm(device=torch.device(f"cuda:{local_rank}"))
    dtype = "complex128" if args.dtype == "c128" else "complex64"
    n = max(1, case.n_qubits)
    # the scaling family keeps 2^(n - log2 N) amplitudes per GPU; its N = 1 member is measured here on
    # rank 0 (bench.py's own N = 1 run is BASELINE config 2, a different circuit), so that the line
    # carries what a weak-scaling efficiency has to be computed against
    family_n1 = None
    if rank == 0 and getattr(args, "family_n1", None) is not None:
        family_n1 = _family_single_gpu(args.family_n1, dtype, local_rank, args.jit)
    comm.barrier()
    sv = ShardedStateVector(n, comm, make_gpu_shard_factory(dtype, local_rank))
    exchange = "peer-memory pull (CUDA IPC mapped send buffers, one NVLink copy per swap)"
    if args.exchange == "nccl" or not comm.enable_peer_pull(sv.shard.send):
        exchange = "NCCL grouped send/recv, 8-piece pack/exchange/unpack pipeline"
    prog = SubspaceProgram(case.meta["subspace_out"])
    norm = case.meta["normalization"]
    opts = PlanOpts.make(timing=True, jit=args.jit)
    rng = np.random.default_rng(0)
    dim_in = SubspaceProgram(case.meta["subspace_in"]).dimension
    inputs = [int(i) for i in rng.integers(0, min(dim_in, 1 << 20), size=args.steps + args.warmup)]
    live = [g for g in case.gates if g.name.upper() != "I"]

    def step(j):
        sv.set_basis(j)
        sv.apply(live, opts)
        return abs(norm) * math.sqrt(sv.project_norm2(prog))

    for j in inputs[: args.warmup]:
        step(j)
    clocks = ClockSampler(index=local_rank)
    if rank == 0:
        clocks.start()
    for k in sv.stats:
        sv.stats[k] = 0
    sv.shard.sv.stats_reset()
    comm.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sv.shard.sv.mark(0)
    norms = [step(j) for j in inputs[args.warmup :]]
    sv.shard.sv.mark(1)
    torch.cuda.synchronize()
    comm.barrier()
    wall_s = time.perf_counter() - t0
    # every step ends in an all-reduce + host read, so the engine-stream event span is the rank's time
    ms_rank = sv.shard.sv.elapsed_ms(0, 1)
    ms = comm.allreduce_max(ms_rank)
    st = sv.shard.sv.stats()
    swap_s = comm.allreduce_max(sv.stats["swap_s"])
    if rank == 0:
        clk = clocks.stop()
        gates_live = len(live)
        peak, peak_src = measured_peak()
        cls = max(("tile", "low6", "perm", "perm_jit", "gate1q"), key=lambda c: st["ms_by_class"][c])
        L = max(1, st["launches_by_class"][cls])
        per_ms = st["ms_by_class"][cls] / L
        per_bytes = st["alg_bytes_by_class"][cls] / L
        achieved = per_bytes / (per_ms * 1e-3) / 1e9 if per_ms > 0 else 0.0
        swap_bytes = sv.stats["swap_bytes"]
        line = {
            "metric": "gates_per_sec",
            "value": gates_live * args.steps / (ms * 1e-3),
            "unit": "gates/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": args.dtype,
            "data": "synthetic",
            "config": config_dict(case, world, args.dtype),
            "gates_live_per_step": gates_live,
            "roofline": {
                "bound": "hbm",
                "kernel": cls,
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "peak_source": peak_src,
                "alg_bytes_per_launch": per_bytes,
                "ms_per_launch": per_ms,
                "launches": st["launches_by_class"][cls],
                "traffic": None,
            },
            "exchange": {
                "swaps_per_step": sv.stats["swaps"] / args.steps,
                "bytes_sent_per_rank_per_step": swap_bytes / args.steps,
                "seconds_per_step_max_rank": swap_s / args.steps,
                "achieved_GBps_per_direction": (swap_bytes / swap_s / 1e9) if swap_s > 0 else None,
                "nvlink_peak_GBps_per_direction": 770.0,
                "path": exchange,
                "breakdown_s_per_step_rank0": {k: sv.stats[k] / args.steps for k in ("pack_s", "barrier_s", "transfer_s", "unpack_s") if k in sv.stats},
                "note": "seconds include pack + transfer + unpack",
            },
            "weak_scaling_n1": family_n1,
            "kernels_rank0": {c: {"launches": st["launches_by_class"][c], "ms": st["ms_by_class"][c]} for c in st["ms_by_class"] if st["launches_by_class"][c]},
            "cpu_baseline": None,
            "e2e": {
                "value": gates_live * args.steps / wall_s,
                "unit": "gates/s",
                "h2d_bytes_per_step": 80 * gates_live,
                "d2h_bytes_per_step": 8,
                "note": "wall clock around the same steps: ShardedStateVector.set_basis/apply/project_norm2 from host gate lists; the result is the scalar norm",
            },
            "gpu_launches": int(st["launches"]) * world,
            "clocks": clk,
            "check": {"norm_first": float(norms[0]), "norm_last": float(norms[-1])},
        }
        print(json.dumps(line), flush=True)
    sv.close()
    dist.destroy_process_group()

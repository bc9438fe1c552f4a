#!/usr/bin/env python
"""bench.py -- gates/sec of the simulate hot path on BASELINE.json's configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One *step* = one pass of the hot path over one input: load basis state |j>, apply the whole block-
encoding circuit, reduce the post-selected norm (what ``Node.simulate_norm(j)`` does,
/root/reference/src/unitaria/nodes/node.py:390-403).  N=1 runs BASELINE config 2 (1-D Laplace block
encoding, bits=26: 29 qubits, 1450 tequila gates, complex128, 8 GiB state); the gate list is the
reference's own ``node.circuit()`` output, frozen in tests/golden/ (the reference does not exist on
the GPU box).  N>1 runs config 5's family sharded over the GPUs by global qubits (weak scaling:
2^31 amplitudes per GPU).  One JSON line is printed by rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
NVLINK_GBS = 770.0         # measured peer copy per direction (same guide)


def load_workload(name):
    from unitaria_b200.fixtures import load_circuit

    return load_circuit(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel_class: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(kernel_class)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index
        self.window = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if self.window is None or self.window[0] <= t <= self.window[1]]
        scope = "timed region"
        if len(rows) < 2:  # timed region shorter than the sampling period: use every sample under load
            rows, scope = [r for _t, r in self.rows], "warm-up + timed region"
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference-algorithm CPU restatement (oracle/sv_oracle.c)
# ---------------------------------------------------------------------------------------------


class CpuState:
    """The CPU port's full-width complex128 register, allocated once (8 GiB for 29 qubits)."""

    def __init__(self, case):
        from oracle import c_oracle

        self.n = max(1, case.n_qubits)
        while True:
            try:
                self.psi = np.empty(2**self.n, dtype=np.complex128)
                break
            except MemoryError:
                self.n -= 1  # host RAM too small for the full register: shrink and say so
        c_oracle.lib().orc_set_basis(self.psi.ctypes.data, self.n, 0)
        live = [g for g in case.gates if g.name.upper() != "I"]
        self.live = len(live)
        self.keep = [g for g in live if all(q < self.n for q in tuple(g.target) + tuple(g.control))]


def cpu_sample(case, budget_s=20.0, stride=4, state=None):
    """Time the CPU port on a bounded sample of the workload: every `stride`-th live gate of the
    circuit applied to the FULL-width complex128 state (per-gate cost does not depend on the
    amplitudes' values).  Returns gates/s over the sample."""
    from oracle import c_oracle

    st = state or CpuState(case)
    n, psi = st.n, st.psi
    sample = st.keep[::stride]
    enc = c_oracle.encode_gates(sample)
    c_oracle.apply_encoded(psi, n, enc[:2])  # touch pages / warm up
    done, t0 = 0, time.perf_counter()
    chunk = 8
    while done < len(enc):
        c_oracle.apply_encoded(psi, n, enc[done : done + chunk])
        done += min(chunk, len(enc) - done)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {
        "value": done / dt,
        "unit": "gates/s",
        "cores": c_oracle.max_threads(),
        "kind": "port",
        "sample": f"{done} gates (every {stride}th live gate of {st.live}) on the full {n}-qubit complex128 state, "
        f"{dt:.1f} s; reference-algorithm CPU restatement (one OpenMP sweep per gate, no fusion), not qulacs",
        "seconds": dt,
        "qubits": n,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = load_workload(args.workload or workload_for(args.gpus))
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    state = CpuState(case)
    for _ in range(args.warmup):
        cpu_sample(case, budget_s=per_step_budget / 2, state=state)
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_sample(case, budget_s=per_step_budget, state=state)
        vals.append(last["value"])
    total = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference",
        "metric": "gates_per_sec",
        "value": value,
        "unit": "gates/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "c128",
        "data": "synthetic",
        "config": config_dict(case, args.gpus, "c128"),
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")} | {"value": value},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------


def workload_for(n_gpus: int) -> str:
    return {1: "cfg2_laplace_n26", 2: "cfg5_bits25", 4: "cfg5_bits26", 8: "cfg5_bits27"}[n_gpus]


def config_dict(case, n_gpus, dtype):
    return {
        "workload": f"{case.meta['name']}: {case.meta['n_qubits']} qubits, {case.meta['n_gates']} tequila gates "
        f"(reference node.circuit() output, frozen in tests/golden/), basis-state inputs, step = set_basis + circuit + post-selected norm",
        "qubits": case.meta["n_qubits"],
        "gates": case.meta["n_gates"],
        "state_bytes": (16 if dtype == "c128" else 8) << max(1, case.meta["n_qubits"]),
        "parallelism": "single GPU" if n_gpus == 1 else f"{n_gpus} GPUs, statevector sharded by the top {int(np.log2(n_gpus))} qubits",
        "l2": "every sweep streams the whole state (>= 8 GiB), far beyond the 126 MB L2; no flush between steps",
    }


def live_gate_count(case) -> int:
    from unitaria_b200 import cabi
    from unitaria_b200.lowering import lower_gates

    low = lower_gates(case.gates, case.n_qubits)
    segs, _ = cabi.plan_describe(max(1, case.n_qubits), 0, low, cabi.PlanOpts.make())
    return sum(len(idx) for _c, _m, idx in segs)


def run_single(args):
    from unitaria_b200 import PlanOpts, StateVector, SubspaceProgram, adapter, cabi, lower_gates

    if cabi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: unitaria_b200 has no CPU fallback")
    case = load_workload(args.workload or workload_for(1))
    dtype = args.dtype
    n = max(1, case.n_qubits, args.qubits or 0)  # --qubits pads the register with idle high qubits (A (x) Identity)
    norm = case.meta["normalization"]
    prog = SubspaceProgram(case.meta["subspace_out"])
    prog_in = SubspaceProgram(case.meta["subspace_in"])
    lowered = lower_gates(case.gates, n)
    amp_bytes = 16 if dtype == "c128" else 8
    rng = np.random.default_rng(0)
    # permutation runs: compile the NVRTC-specialised kernel in the first warm-up step instead of in
    # the background, so that every timed step runs the same kernels
    opts = PlanOpts.make(timing=True, jit=args.jit, mux=not args.no_mux, phases=not args.no_phases, tile_qubits=args.tile_qubits, lookahead=args.lookahead, tile_io=args.tile_io)

    sv = StateVector(n, "complex128" if dtype == "c128" else "complex64")
    n_inputs = args.steps + args.warmup
    if prog_in.total_qubits <= 24:
        basis_in = sv.enumerate_basis(prog_in)
        inputs = [int(basis_in[i]) for i in rng.integers(0, len(basis_in), size=n_inputs)]
    else:  # big input subspaces: members by rank (the inverse of the device's rank computation), no enumeration
        from unitaria_b200.subspace_prog import spec_member, to_spec

        spec_in = to_spec(case.meta["subspace_in"])
        inputs = [spec_member(spec_in, int(r)) for r in rng.integers(0, prog_in.dimension, size=n_inputs)]

    def step(j):
        sv.set_basis(j)
        sv.apply_lowered(lowered, opts)
        return abs(norm) * np.sqrt(sv.project_norm2(prog))

    clocks = ClockSampler()
    clocks.start()
    for j in inputs[: args.warmup]:
        step(j)
    sv.stats_reset()
    sv.sync()
    t_begin = time.perf_counter()
    sv.mark(0)
    norms = [step(j) for j in inputs[args.warmup :]]
    sv.mark(1)
    sv.sync()
    clocks.window = (t_begin, time.perf_counter())
    ms = sv.elapsed_ms(0, 1)
    st = sv.stats()
    gates_live = st["gates_live"] // max(1, args.steps)
    value = gates_live * args.steps / (ms * 1e-3)

    # dominant kernel class by device time
    cls = max(("tile", "perm", "perm_jit", "gate1q", "low6"), key=lambda c: st["ms_by_class"][c])
    launches = st["launches_by_class"][cls]
    peak, peak_src = measured_peak()
    per_launch_bytes = st["alg_bytes_by_class"][cls] / max(1, launches)
    per_launch_ms = st["ms_by_class"][cls] / max(1, launches)
    achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    roofline = {
        "bound": "hbm",
        "kernel": {
            "tile": "k_tile (fused tile sweep)",
            "perm": "k_perm (permutation-run sweep, generic)",
            "perm_jit": "b2sv_perm_jit (permutation-run sweep, NVRTC-specialised)",
            "gate1q": "k_gate1",
            "low6": "k_low6 (low-qubit warp-shuffle sweep)",
        }[cls],
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "peak_source": peak_src,
        "alg_bytes_per_launch": per_launch_bytes,
        "ms_per_launch": per_launch_ms,
        "launches": launches,
        "traffic": recorded_traffic(cls),
        "note": "per launch = total algorithmic bytes / total device time over the launches of the step.  After set_basis the "
        "first sweep runs only on the tile that owns the basis index when a permutation sweep follows (live window: 64 KiB, "
        "and that permutation sweep is write-only: 1*B*2^n), or generates the basis state in shared memory and only WRITES "
        "the state otherwise (1*B*2^n, counted as such); write-only traffic runs above the read+write copy peak",
    }
    kernels = {}
    for c in ("tile", "low6", "perm", "perm_jit", "gate1q", "project", "init", "scale"):
        L = st["launches_by_class"][c]
        if L:
            t = st["ms_by_class"][c]
            b = st["alg_bytes_by_class"][c]
            kernels[c] = {
                "launches_per_step": L / args.steps,
                "ms_per_step": t / args.steps,
                "alg_GBps": (b / (t * 1e-3) / 1e9) if t > 0 else None,
                "frac_of_peak": (b / (t * 1e-3) / 1e9 / peak) if t > 0 else None,
            }
    clk = clocks.stop()
    sv.close()

    # ---- end to end through the public API with host buffers (Node.simulate(j) fast path) ----
    class Circuit:
        gates = case.gates

    adapter.configure(dtype="complex128" if dtype == "c128" else "complex64", opts=PlanOpts.make(jit=args.jit))
    e2e_steps = max(1, min(args.steps, 20))
    e2e = None
    if not args.no_e2e:
        for j in inputs[: min(3, len(inputs))]:  # warm: lowering cache, resident register, pinned result buffers
            out = adapter.simulate_projected(Circuit, n, prog, norm, j)
        t0 = time.perf_counter()
        for j in inputs[args.warmup : args.warmup + e2e_steps]:
            out = adapter.simulate_projected(Circuit, n, prog, norm, j)
        e2e_s = time.perf_counter() - t0
        e2e = {
            "value": gates_live * e2e_steps / e2e_s,
            "unit": "gates/s",
            "h2d_bytes_per_step": int(lowered.nbytes + 96 * len(lowered)),
            "d2h_bytes_per_step": int(out.nbytes),
            "steps": e2e_steps,
            "ms_per_step": 1e3 * e2e_s / e2e_steps,
            "call": "unitaria_b200.simulate_projected(circuit, n_qubits, subspace_out, normalization, j) -> host complex128[dim_out] "
            "(= Node.simulate(j)); H2D = lowered gate program, D2H = projected vector",
        }
    adapter.clear_caches()

    cpu = None if args.no_cpu_baseline else cpu_sample(case, budget_s=args.cpu_seconds)
    line = {
        "metric": "gates_per_sec",
        "value": value,
        "unit": "gates/s",
        "n_gpus": 1,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": dtype,
        "data": "synthetic",
        "config": config_dict(case, 1, dtype),
        "gates_live_per_step": gates_live,
        "roofline": roofline,
        "kernels": kernels,
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": e2e,
        "gpu_launches": int(st["launches"]),
        "clocks": clk,
        "check": {"norm_first": float(norms[0]), "norm_last": float(norms[-1])},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="fixture name under tests/golden (default: BASELINE config for --gpus)")
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    ap.add_argument("--qubits", type=int, default=None, help="pad the register to this many qubits (idle high qubits)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mux", action="store_true", help="A/B switch: disable same-target multiplexor fusion")
    ap.add_argument("--no-phases", action="store_true", help="A/B switch: disable register phases in tile sweeps")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU swap transport")
    ap.add_argument("--tile-qubits", type=int, default=0)
    ap.add_argument("--tile-io", type=int, default=0, help="0 one tile per CTA with bulk-async copies (default), 1 plain ld/st, 3 experimental persistent pipeline")
    ap.add_argument("--lookahead", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--jit", default="sync", choices=["sync", "auto", "off"], help="permutation-run kernel specialisation")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    args.family_n1 = None
    if args.warmup < 3 and args.impl == "ours" and not args.workload:
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.gpus == 1:
        run_single(args)
    else:
        from unitaria_b200 import distributed

        if not args.workload:  # N = 1 member of the scaling family: same amplitudes per GPU
            args.family_n1 = load_workload("cfg5_bits24")
        distributed.bench_main(args, load_workload(args.workload or workload_for(args.gpus)), config_dict, cpu_sample, ClockSampler, measured_peak)


if __name__ == "__main__":
    main()

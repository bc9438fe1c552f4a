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
    """The CPU port's complex128 register, allocated once.  The sample runs on the FULL register when it has at
    most ``max_qubits`` qubits (29 qubits = 8 GiB for config 2) and on a ``max_qubits``-qubit register otherwise:
    one in-place sweep per gate costs time proportional to the amplitudes it visits, so the full-size rate is the
    sampled rate times 2^(sample qubits - full qubits) -- stated in the line's ``sample`` text."""

    def __init__(self, case, max_qubits=30):
        from oracle import c_oracle

        self.n_full = max(1, case.n_qubits)
        self.n = min(self.n_full, max_qubits)
        try:
            import psutil

            avail = psutil.virtual_memory().available
            while self.n > 20 and (16 << self.n) > 0.6 * avail:
                self.n -= 1
        except Exception:
            pass
        self.psi = np.zeros(2**self.n, dtype=np.complex128)  # zeros: touches every page now, not inside the timing
        c_oracle.lib().orc_set_basis(self.psi.ctypes.data, self.n, 0)
        live = [g for g in case.gates if g.name.upper() != "I"]
        self.live = len(live)
        # gates that do not fit the sampled register are re-homed on it (same control count, same sweep size)
        shift = self.n_full - self.n
        self.keep = live if shift == 0 else [_fold_gate(g, self.n) for g in live]
        self.scale = 2.0 ** (self.n - self.n_full)


def _fold_gate(g, n):
    """The same gate on qubits folded into an n-qubit register (distinct qubits stay distinct): per-gate CPU cost
    depends on the number of controls and the register size, not on which qubits they are."""
    from unitaria_b200.fixtures import GateRecord

    qs = [int(q) for q in tuple(g.target or ()) + tuple(g.control or ())]
    used, out = set(), {}
    for q in qs:
        f = q % n
        while f in used:
            f = (f + 1) % n
        used.add(f)
        out[q] = f
    return GateRecord(g.name, tuple(out[int(t)] for t in (g.target or ())), tuple(out[int(c)] for c in (g.control or ())), g.parameter, getattr(g, "power", 1.0))


def cpu_sample(case, budget_s=20.0, stride=4, state=None, max_qubits=30):
    """Time the CPU port on a bounded sample of the workload: every `stride`-th live gate of the
    circuit applied to a complex128 state (per-gate cost does not depend on the
    amplitudes' values).  Returns gates/s over the sample, scaled to the full register size."""
    from oracle import c_oracle

    st = state or CpuState(case, max_qubits)
    n, psi = st.n, st.psi
    sample = st.keep[::stride]
    enc = c_oracle.encode_gates(sample)
    c_oracle.apply_encoded(psi, n, enc[:2])  # warm up
    done, t0 = 0, time.perf_counter()
    chunk = 8
    while done < len(enc):
        c_oracle.apply_encoded(psi, n, enc[done : done + chunk])
        done += min(chunk, len(enc) - done)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    scaled = "" if st.n == st.n_full else f"; sampled on {n} qubits and scaled by 2^({n}-{st.n_full}) to the full {st.n_full}-qubit register (one sweep per gate: time ~ amplitudes visited)"
    return {
        "value": done / dt * st.scale,
        "unit": "gates/s",
        "cores": c_oracle.max_threads(),
        "kind": "port",
        "sample": f"{done} gates (every {stride}th live gate of {st.live}) on a {n}-qubit complex128 state, "
        f"{dt:.1f} s; reference-algorithm CPU restatement (one OpenMP sweep per gate, no fusion), not qulacs" + scaled,
        "seconds": dt,
        "qubits": n,
    }


def effective_bytes(lowered, n, amp_bytes):
    """SURVEY.md 8(d) "effective" traffic: what a gate-at-a-time simulator would move for this circuit,
    sum over lowered gates of 2*B*2^(n-c) (c controls; half of that for a diagonal gate with one unit entry, which only
    touches target = 1).  Fusion makes effective GB/s exceed the HBM peak: it is NOT a roofline fraction."""
    from unitaria_b200 import cabi

    c = np.array([bin(int(m)).count("1") for m in lowered["ctrl_mask"]], dtype=np.float64)
    half = (lowered["op"] == cabi.OP_DIAG1) & (lowered["m"][:, 0] == 1.0) & (lowered["m"][:, 1] == 0.0)
    c = c + half.astype(np.float64)
    return float(np.sum(2.0 * amp_bytes * np.exp2(n - c)))


def reference_projection_cost(spec, seconds=2.0):
    """The reference's own projection (subspace.py:325-335) filters range(2^total_qubits) through the Python predicate
    test_basis, one index at a time; this times the oracle's literal restatement of that predicate on a bounded sample
    of indices (part of today's Node.simulate, reported apart from the gate sweeps as SURVEY.md 8(d) asks)."""
    from oracle import np_oracle

    total = np_oracle.spec_total_qubits(spec)
    rng = np.random.default_rng(0)
    idx = [int(i) for i in rng.integers(0, 1 << total, size=200000)]
    done, t0 = 0, time.perf_counter()
    while done < len(idx) and time.perf_counter() - t0 < seconds:
        for i in idx[done : done + 2000]:
            np_oracle.test_basis(spec, i)
        done += 2000
    dt = time.perf_counter() - t0
    per = dt / max(1, done)
    return {
        "us_per_index": 1e6 * per,
        "indices_sampled": done,
        "total_qubits": total,
        "enumerate_basis_seconds_at_full_size": per * float(1 << total),
        "note": "oracle/np_oracle.test_basis (restates subspace.py:297-323) on random indices, one host thread; the reference calls it for every index of range(2^total_qubits) on each projection",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle

    c_oracle.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; this arm is the box's host cores
    case = load_workload(args.workload or workload_for(args.gpus))
    # the whole arm stays under ~2 minutes whatever N is: at N > 1 (32-34 qubit registers) the sample runs on 30 qubits
    per_step_budget = max(1.5, min(20.0, 100.0 / max(1, args.steps + args.warmup)))
    state = CpuState(case, max_qubits=30)
    for _ in range(min(args.warmup, 2)):
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


KERNEL_NAMES = {
    "tile": "k_tile_swz (fused tile sweep, TMA tensor copies)",
    "perm": "k_perm (permutation-run sweep, generic)",
    "perm_jit": "b2sv_perm_jit (permutation-run sweep, NVRTC-specialised)",
    "gate1q": "k_gate1",
    "low6": "k_low6 (low-qubit warp-shuffle sweep)",
    "dist_perm": "b2sv_perm_dist (distributed permutation sweep: gather over peer-mapped shards)",
}


def pick_inputs(case, sv, rng, count):
    """Basis inputs of the encoded matrix: members of subspace_in, picked by rank."""
    from unitaria_b200 import SubspaceProgram
    from unitaria_b200.subspace_prog import spec_member, to_spec

    prog_in = SubspaceProgram(case.meta["subspace_in"])
    spec_in = to_spec(case.meta["subspace_in"])
    return [spec_member(spec_in, int(r)) for r in rng.integers(0, prog_in.dimension, size=count)], prog_in


def measure(case, dtype, n, steps, warmup, opts, dense=False, seed=0):
    """`steps` timed steps of one workload on one GPU: load the input (a basis state |j>, or -- dense -- a seeded
    dense vector on the input subspace's qubits, generated on the device and NOT timed), apply the whole circuit,
    reduce the post-selected norm.  Device time by CUDA events on the engine's stream."""
    from unitaria_b200 import StateVector, SubspaceProgram, lower_gates

    norm = case.meta["normalization"]
    prog = SubspaceProgram(case.meta["subspace_out"])
    lowered = lower_gates(case.gates, n)
    rng = np.random.default_rng(seed)
    peak, _src = measured_peak()
    sv = StateVector(n, "complex128" if dtype == "c128" else "complex64")
    try:
        inputs, prog_in = pick_inputs(case, sv, rng, steps + warmup)
        support = prog_in.total_qubits

        def step(k, j, timed):
            if dense:
                sv.fill_random(seed=1000 + k, support_bits=support)
            else:
                sv.set_basis(j)
            if timed:
                sv.mark(2)
            sv.apply_lowered(lowered, opts)
            r = abs(norm) * np.sqrt(sv.project_norm2(prog))
            if timed:
                sv.mark(3)
            return r

        for k, j in enumerate(inputs[:warmup]):
            step(k, j, False)
        sv.sync()
        sv.stats_reset()
        t_begin = time.perf_counter()
        ms, norms = 0.0, []
        if dense:  # the fill is not part of the step: time each step between its own marks
            for k, j in enumerate(inputs[warmup:]):
                norms.append(step(k, j, True))
                ms += sv.elapsed_ms(2, 3)
        else:
            sv.mark(0)
            norms = [step(k, j, False) for k, j in enumerate(inputs[warmup:])]
            sv.mark(1)
            sv.sync()
            ms = sv.elapsed_ms(0, 1)
        window = (t_begin, time.perf_counter())
        st = sv.stats()
    finally:
        sv.close()
    gates_live = st["gates_live"] // max(1, steps)
    kernels = {}
    for c in ("tile", "low6", "perm", "perm_jit", "gate1q", "project", "init", "scale"):
        L = st["launches_by_class"][c]
        if L and not (dense and c == "init"):
            t = st["ms_by_class"][c]
            bts = st["alg_bytes_by_class"][c]
            kernels[c] = {
                "launches_per_step": L / steps,
                "ms_per_step": t / steps,
                "alg_GBps": (bts / (t * 1e-3) / 1e9) if t > 0 else None,
                "frac_of_peak": (bts / (t * 1e-3) / 1e9 / peak) if t > 0 else None,
            }
    return {
        "ms": ms,
        "ms_per_step": ms / steps,
        "gates_live": int(gates_live),
        "value": gates_live * steps / (ms * 1e-3),
        "kernels": kernels,
        "stats": st,
        "norms": norms,
        "window": window,
        "lowered": lowered,
        "inputs": inputs,
        "prog": prog,
    }


def roofline_of(m, steps, dense_note=""):
    st = m["stats"]
    cls = max(("tile", "perm", "perm_jit", "gate1q", "low6"), key=lambda c: st["ms_by_class"][c])
    launches = st["launches_by_class"][cls]
    peak, peak_src = measured_peak()
    per_launch_bytes = st["alg_bytes_by_class"][cls] / max(1, launches)
    per_launch_ms = st["ms_by_class"][cls] / max(1, launches)
    achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    return {
        "bound": "hbm",
        "kernel": KERNEL_NAMES[cls],
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "peak_source": peak_src,
        "alg_bytes_per_launch": per_launch_bytes,
        "ms_per_launch": per_launch_ms,
        "launches": launches,
        "traffic": recorded_traffic(cls),
        "note": "per launch = total algorithmic bytes / total device time over the launches of the class in the timed steps"
        + dense_note,
    }


def summarise(m, steps):
    r = roofline_of(m, steps)
    return {
        "ms_per_step": m["ms_per_step"],
        "gates_per_sec": m["value"],
        "gates_live_per_step": m["gates_live"],
        "steps": steps,
        "dominant_kernel": r["kernel"],
        "frac": r["frac"],
        "kernels": m["kernels"],
        "norm_first": float(m["norms"][0]),
    }


# BASELINE.json configs 3 and 4 next to the headline (config 2): (fixture, dtype, register qubits, steps, warm-up)
EXTRA_CONFIGS = {
    "cfg3_gaussian_28q_c64": ("cfg3_gaussian_s3_t21", "c64", 28, 20, 3),
    "cfg4_bpx_L8_30q": ("cfg4_bpx_L8", "c128", 30, 5, 3),
    "cfg4_pinv_bpx_L4_30q": ("cfg4_pinv_bpx_L4", "c128", 30, 1, 1),
}


def run_single(args):
    from unitaria_b200 import PlanOpts, adapter, cabi

    if cabi.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: unitaria_b200 has no CPU fallback")
    case = load_workload(args.workload or workload_for(1))
    dtype = args.dtype
    n = max(1, case.n_qubits, args.qubits or 0)  # --qubits pads the register with idle high qubits (A (x) Identity)
    norm = case.meta["normalization"]
    # permutation runs: compile the NVRTC-specialised kernel in the first warm-up step instead of in
    # the background, so that every timed step runs the same kernels
    opts = PlanOpts.make(timing=True, jit=args.jit, mux=not args.no_mux, phases=not args.no_phases, tile_qubits=args.tile_qubits, lookahead=args.lookahead, tile_io=args.tile_io)

    clocks = ClockSampler()
    clocks.start()
    m = measure(case, dtype, n, args.steps, args.warmup, opts)
    clocks.window = m["window"]
    clk = clocks.stop()
    st = m["stats"]
    gates_live = m["gates_live"]
    roofline = roofline_of(
        m,
        args.steps,
        ".  After set_basis the first sweep runs only on the tile that owns the basis index when a permutation sweep follows "
        "(live window: 64 KiB, and that permutation sweep is write-only: 1*B*2^n), or generates the basis state in shared memory and "
        "only WRITES the state otherwise (1*B*2^n, counted as such); write-only traffic runs above the read+write copy peak.  "
        "`inputs.dense` times the same circuit on a dense input, where every sweep reads and writes the whole register",
    )
    # the same workload on a seeded dense vector embedded in the input subspace's qubits: no lazy
    # initialisation, no live window -- every sweep is a full read+write
    inputs = {"basis": {"ms_per_step": m["ms_per_step"], "gates_per_sec": m["value"], "kernels": m["kernels"]}}
    if not args.no_dense:
        dsteps = max(1, min(args.steps, 10))
        d = measure(case, dtype, n, dsteps, min(args.warmup, 3), opts, dense=True)
        inputs["dense"] = summarise(d, dsteps)
        inputs["dense"]["input"] = "b2sv_fill_random: seeded dense vector on the input subspace's qubits, generated on the device outside the timed region"
    configs = {}
    if not args.workload and not args.no_extra_configs:
        for key, (fixture, cdtype, cn, csteps, cwarm) in EXTRA_CONFIGS.items():
            try:
                ccase = load_workload(fixture)
                cm = measure(ccase, cdtype, max(cn, ccase.n_qubits), csteps, cwarm, PlanOpts.make(timing=True, jit="sync"))
                configs[key] = summarise(cm, csteps) | {"qubits": max(cn, ccase.n_qubits), "dtype": cdtype, "tequila_gates": ccase.meta["n_gates"]}
            except Exception as exc:  # noqa: BLE001 - an extra config must not take the headline down with it
                configs[key] = {"error": repr(exc)}

    # ---- end to end through the public API with host buffers (Node.simulate(j) fast path) ----
    class Circuit:
        gates = case.gates

    adapter.configure(dtype="complex128" if dtype == "c128" else "complex64", opts=PlanOpts.make(jit=args.jit))
    e2e_steps = max(1, min(args.steps, 20))
    e2e = None
    lowered, prog, in_list = m["lowered"], m["prog"], m["inputs"]
    if not args.no_e2e:
        for j in in_list[: min(3, len(in_list))]:  # warm: lowering cache, resident register, pinned result buffers
            out = adapter.simulate_projected(Circuit, n, prog, norm, j)
        t0 = time.perf_counter()
        for j in in_list[args.warmup : args.warmup + e2e_steps]:
            out = adapter.simulate_projected(Circuit, n, prog, norm, j)
        e2e_s = time.perf_counter() - t0
        sv_e2e = adapter._state(n)
        d2h = int(getattr(sv_e2e, "last_d2h_bytes", out.nbytes))
        e2e = {
            "value": gates_live * e2e_steps / e2e_s,
            "unit": "gates/s",
            "h2d_bytes_per_step": int(lowered.nbytes + 96 * len(lowered)),
            "d2h_bytes_per_step": d2h,
            "result_bytes_on_host": int(out.nbytes),
            "steps": e2e_steps,
            "ms_per_step": 1e3 * e2e_s / e2e_steps,
            "call": "unitaria_b200.simulate_projected(circuit, n_qubits, subspace_out, normalization, j) -> host complex128[dim_out] "
            "(= Node.simulate(j)); H2D = lowered gate program, D2H = the projected vector -- as (rank, value) pairs of its non-zeros "
            "when they are few (b2sv_project_gather_sparse; the dense array is assembled on the host), else the whole array",
        }
        # the same call with the compact form switched off: the whole dim_out array crosses PCIe (what a dense column costs)
        saved_min = type(sv_e2e).SPARSE_MIN_DIM
        type(sv_e2e).SPARSE_MIN_DIM = 1 << 62
        try:
            for j in in_list[:3]:  # warm: two page-locked 1 GiB result buffers end up in the pool
                out = adapter.simulate_projected(Circuit, n, prog, norm, j)
            t0 = time.perf_counter()
            for j in in_list[args.warmup : args.warmup + e2e_steps]:
                out = adapter.simulate_projected(Circuit, n, prog, norm, j)
            dense_s = time.perf_counter() - t0
            e2e["dense_copy"] = {
                "value": gates_live * e2e_steps / dense_s,
                "unit": "gates/s",
                "ms_per_step": 1e3 * dense_s / e2e_steps,
                "d2h_bytes_per_step": int(out.nbytes),
            }
        finally:
            type(sv_e2e).SPARSE_MIN_DIM = saved_min
        # matrix case of Node.simulate (node.py:381-388): columns pipelined, the D2H of column j overlaps column j+1
        if hasattr(adapter, "simulate_columns"):
            cols = in_list[args.warmup : args.warmup + e2e_steps]
            adapter.simulate_columns(Circuit, n, prog, norm, cols[:2])
            t0 = time.perf_counter()
            adapter.simulate_columns(Circuit, n, prog, norm, cols, keep=False)
            pipe_s = time.perf_counter() - t0
            e2e["pipelined_columns"] = {
                "value": gates_live * len(cols) / pipe_s,
                "unit": "gates/s",
                "ms_per_column": 1e3 * pipe_s / len(cols),
                "columns": len(cols),
                "call": "unitaria_b200.simulate_columns(...): same host result per column; compact form when the column is sparse, else double-buffered D2H on a second stream",
            }
    adapter.clear_caches()

    cpu = None if args.no_cpu_baseline else cpu_sample(case, budget_s=args.cpu_seconds)
    if cpu is not None:
        cpu["projection"] = reference_projection_cost(case.meta["subspace_out"])
    eff_bytes = effective_bytes(lowered, n, 16 if dtype == "c128" else 8)
    line = {
        "metric": "gates_per_sec",
        "value": m["value"],
        "unit": "gates/s",
        "n_gpus": 1,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": dtype,
        "data": "synthetic",
        "config": config_dict(case, 1, dtype),
        "gates_live_per_step": gates_live,
        "roofline": roofline,
        "kernels": m["kernels"],
        "inputs": inputs,
        "configs": configs,
        "effective": {
            "GBps": eff_bytes / (m["ms_per_step"] * 1e-3) / 1e9,
            "bytes_per_step": eff_bytes,
            "note": "sum over lowered gates of 2*B*2^(n-controls) / device time per step: the traffic of a gate-at-a-time simulator; "
            "above the HBM peak because sweeps are fused -- an 'effective' figure, not a roofline fraction",
        },
        "cpu_baseline": None if cpu is None else {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "projection")},
        "e2e": e2e,
        "gpu_launches": int(st["launches"]),
        "clocks": clk,
        "check": {"norm_first": float(m["norms"][0]), "norm_last": float(m["norms"][-1])},
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
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-input timing of the headline workload")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip BASELINE configs 3 and 4 (cfg3 c64, BPX and Pseudoinverse(BPX) at 30 qubits)")
    ap.add_argument("--no-mux", action="store_true", help="A/B switch: disable same-target multiplexor fusion")
    ap.add_argument("--no-phases", action="store_true", help="A/B switch: disable register phases in tile sweeps")
    ap.add_argument("--exchange", default="gather", choices=["gather", "nccl"], help="multi-GPU exchange: fused gather over peer-mapped shards (default) or pack / NCCL send-recv / unpack")
    ap.add_argument("--tile-qubits", type=int, default=0)
    ap.add_argument("--tile-io", type=int, default=0, help="0 one tile per CTA with TMA tensor copies (default), 1 plain ld/st, 2 linear bulk copies")
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

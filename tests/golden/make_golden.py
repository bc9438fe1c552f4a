"""Generate the golden fixtures in this directory from the REFERENCE's own Python.

Run in the build container only (the reference is at /root/reference there and nowhere else):

    python tests/golden/make_golden.py [--big]

It imports the read-only reference package (``/root/reference/src``) against the tequila stand-in
in tests/_tequila_shim (tequila itself is not installable here), builds the reference's nodes with
the reference's own constructors (the ones its tests and examples use), and records for each case

* the tequila-level gate list of ``node.circuit()``             (nodes/node.py:190-227)
* ``n_qubits``, ``normalization``, ``subspace_in/out`` as nested specs
* ``expected`` = ``node.toarray()`` -- the reference's *numpy* ``compute()`` path, which is what
  the reference's own tests compare the simulator against (verifier.py:115-126)

No simulator output is stored as "expected": every expected array comes from the reference's
arithmetic definition of the node, so the fixtures pin simulate-vs-compute parity exactly the way
``ut.verify`` does.  ``--big`` also writes the gate lists of the BASELINE.json configs at full size
(gate lists only).  ``--columns`` freezes, for the mid-size and full-size configs, a handful of COLUMNS of the
encoded matrix -- ``node.compute(e_j)``, the left-hand side of verifier.py:119-126 -- as sparse vectors in
``<name>.cols.npz``, so that the GPU suite can check ``normalization * project(simulate(b_j))`` against the
reference's own numbers at sizes where ``toarray()`` cannot be stored (and ``Adjoint(node)`` / the controlled
variant of verifier.py:137,146-153 for the <= 16-qubit cases: ``<name>__adjoint.npz`` / ``<name>__controlled.npz``).
"""

from __future__ import annotations

import argparse
import importlib.util
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "_tequila_shim"), "/root/reference/src"]
sys.setrecursionlimit(100000)
warnings.simplefilter("ignore")

import unitaria as ut  # noqa: E402

from oracle.np_oracle import spec_from_subspace  # noqa: E402
from unitaria_b200.fixtures import save_circuit  # noqa: E402


def _load_example(name):
    spec = importlib.util.spec_from_file_location(name, f"/root/reference/examples/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def laplace_test_form(n):
    # tests/test_laplace.py:8-12
    C = ut.Increment(bits=n)
    return ut.Add(ut.Scale(ut.Identity(ut.Subspace("#" * n)), -2), ut.Scale(ut.Add(C, ut.Adjoint(C)), 1))


def laplace_readme_form(n):
    # README.md:52-55
    inc = ut.Increment(bits=n)
    return (2 * ut.Identity(dim=2**n) - inc - inc.adjoint())[:-1, :-1]


def gaussian_conv(source_bits, target_bits, out_dim):
    # tests/test_gaussian_conv.py:7-15 generalised in the two register widths
    half = 2 ** (source_bits - 1) - 1
    kernel = np.exp(-((np.arange(-half, half + 1) / 4) ** 2))
    padded_kernel = np.append(kernel, 0)
    prep = ut.Identity(ut.Subspace.from_dim(2**target_bits)) & ut.ConstantVector(np.sqrt(padded_kernel))
    add = ut.IntegerAddition(source_bits=source_bits, target_bits=target_bits)
    const_add = ut.ConstantIntegerAddition(bits=target_bits, constant=-half) & ut.Identity(
        ut.Subspace.from_dim(2**source_bits)
    )
    unprep = ut.Adjoint(prep)
    return (unprep @ const_add @ add @ prep)[:out_dim, :out_dim]


def cfg5(bits, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    return ut.Increment(bits=bits) & ut.ConstantVector(v)


def record(name, node, expected=True, max_expected_entries=1 << 22, extra_meta=None):
    t0 = time.time()
    circuit = node.circuit()
    gates = circuit._tq_circuit.gates
    meta = {
        "name": name,
        "label": type(node).__name__,
        "n_qubits": int(circuit.n_qubits),
        "target_qubits": int(node.target_qubit_count()),
        "clean_ancillae": int(node.clean_ancilla_count()),
        "borrowed_ancillae": int(node.borrowed_ancilla_count()),
        "normalization": float(node.normalization),
        "is_vector": bool(node.is_vector()),
        "dimension_in": int(node.dimension_in),
        "dimension_out": int(node.dimension_out),
        "subspace_in": spec_from_subspace(node.subspace_in),
        "subspace_out": spec_from_subspace(node.subspace_out),
        "n_gates": len(gates),
        "touches_qubits": len(circuit._tq_circuit.qubits) > 0,
        "reference": "tequilahub/unitaria v0.2.0",
    }
    meta.update(extra_meta or {})
    arrays = {}
    if expected and node.dimension_in * node.dimension_out <= max_expected_entries:
        arrays["expected"] = np.asarray(node.toarray(), dtype=np.complex128)
        if not node.is_vector():
            arrays["basis_in"] = np.asarray(node.subspace_in.enumerate_basis(), dtype=np.int64)
        if node.subspace_out.total_qubits <= 22:
            arrays["basis_out"] = np.asarray(node.subspace_out.enumerate_basis(), dtype=np.int64)
    save_circuit(os.path.join(HERE, name + ".npz"), gates, circuit.n_qubits, meta, **arrays)
    print(
        f"{name:34s} q={circuit.n_qubits:2d} gates={len(gates):6d} dim={node.dimension_out}x{node.dimension_in}"
        f" expected={'expected' in arrays}  ({time.time() - t0:.1f}s)"
    )


def record_columns(name, node, n_random=3, seed=0):
    """Columns of the encoded matrix through the reference's numpy path, stored sparse."""
    from unitaria_b200.subspace_prog import spec_member

    t0 = time.time()
    dim_in, dim_out = int(node.dimension_in), int(node.dimension_out)
    spec_in = spec_from_subspace(node.subspace_in)
    rng = np.random.default_rng(seed)
    ranks = sorted(set([0, min(1, dim_in - 1), dim_in // 3, dim_in - 1] + [int(r) for r in rng.integers(0, dim_in, size=n_random)]))
    if node.is_vector():
        ranks = [0]
    idx_all, val_all, ptr = [], [], [0]
    for r in ranks:
        if node.is_vector():
            col = np.asarray(node.compute(), dtype=np.complex128).reshape(-1)
        else:
            e = np.zeros(dim_in, dtype=np.complex128)
            e[r] = 1
            col = np.asarray(node.compute(e), dtype=np.complex128).reshape(-1)
        assert col.shape == (dim_out,), (col.shape, dim_out)
        nz = np.flatnonzero(col != 0)
        idx_all.append(nz.astype(np.int64))
        val_all.append(col[nz])
        ptr.append(ptr[-1] + len(nz))
    basis = np.array([spec_member(spec_in, r) for r in ranks], dtype=np.int64)
    np.savez_compressed(
        os.path.join(HERE, name + ".cols.npz"),
        col_rank=np.array(ranks, dtype=np.int64),
        col_basis=basis,
        col_ptr=np.array(ptr, dtype=np.int64),
        col_index=np.concatenate(idx_all) if idx_all else np.zeros(0, np.int64),
        col_value=np.concatenate(val_all) if val_all else np.zeros(0, np.complex128),
        dim_out=np.int64(dim_out),
    )
    print(f"{name:34s} columns={len(ranks)} nnz={ptr[-1]} dim_out={dim_out}  ({time.time() - t0:.1f}s)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also write the full-size BASELINE.json configs")
    ap.add_argument("--columns", action="store_true", help="freeze columns (node.compute(e_j)) of the mid / full-size configs and the verify legs")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    bpx = _load_example("example_bpx")

    small = {
        # config 1 and its siblings (tests/test_laplace.py)
        "laplace_n1": lambda: laplace_test_form(1),
        "laplace_n2": lambda: laplace_test_form(2),
        "laplace_n3": lambda: laplace_test_form(3),
        "laplace_n4": lambda: laplace_test_form(4),
        "laplace_n8": lambda: laplace_test_form(8),
        "laplace_readme_n4": lambda: laplace_readme_form(4),
        "laplace_readme_n8": lambda: laplace_readme_form(8),
        # tests/nodes/test_constant.py / test_fourier_transform.py / test_qsvt.py / test_classical.py
        "constant_vector4": lambda: ut.ConstantVector(np.array([1, 2j, 1 / 3, -1j / 4])),
        "constant_vector5": lambda: ut.ConstantVector(np.array([0.3, -1.0, 2j, 0.25 - 0.5j, 1.5])),
        "fourier3": lambda: ut.FourierTransform(3),
        "fourier4": lambda: ut.FourierTransform(4),
        "qsvt_inc2_cheb": lambda: ut.QSVT(
            1.23 * ut.Increment(bits=2), np.polynomial.Chebyshev([1, 0, 1], domain=[-1.23, 1.23])
        ),
        "increment4": lambda: ut.Increment(bits=4),
        "integer_addition_2_3": lambda: ut.IntegerAddition(source_bits=2, target_bits=3),
        "const_int_add_4_m3": lambda: ut.ConstantIntegerAddition(bits=4, constant=-3),
        "scale_complex": lambda: ut.Scale(ut.Increment(bits=2), 0.3 - 0.4j),
        # tests/test_gaussian_conv.py (1-D: 11 qubits; 2-D: 16 qubits)
        "gaussian_s3_t4": lambda: gaussian_conv(3, 4, 8),
        "gaussian_s3_t8": lambda: gaussian_conv(3, 8, 128),
        # examples/example_bpx.py
        "bpx_L1": lambda: bpx.preconditioned_half_laplace(1, 1, 2),
        "bpx_L2": lambda: bpx.preconditioned_half_laplace(1, 2, 2),
        "bpx_L3": lambda: bpx.preconditioned_half_laplace(1, 3, 2),
        # config 5 at parity size
        "cfg5_bits5": lambda: cfg5(5),
        "cfg5_bits13": lambda: cfg5(13),
    }

    def sinv_rhs():
        A = bpx.preconditioned_half_laplace(1, 2, 2)
        inv = ut.Pseudoinverse(A, 15, 0.1)
        return inv @ ut.Adjoint(inv) @ ut.ConstantVector(np.ones(A.dimension_in))

    small["pinv_bpx_L2"] = lambda: ut.Pseudoinverse(bpx.preconditioned_half_laplace(1, 2, 2), 15, 0.1)
    small["sinv_rhs_bpx_L2"] = sinv_rhs

    mid = {  # gate lists only; the tests compare the CUDA path with the CPU oracle at these sizes
        "laplace_n12": lambda: laplace_test_form(12),
        "laplace_n16": lambda: laplace_test_form(16),
        "laplace_n20": lambda: laplace_test_form(20),
        "gaussian_s3_t13": lambda: gaussian_conv(3, 13, 2**12),
        "bpx_L4": lambda: bpx.preconditioned_half_laplace(1, 4, 2),
        "bpx_L5": lambda: bpx.preconditioned_half_laplace(1, 5, 2),
    }
    big = {  # BASELINE.json configs 2-5 at full size
        "cfg2_laplace_n26": lambda: laplace_test_form(26),
        "cfg2_laplace_readme_n26": lambda: laplace_readme_form(26),
        "cfg3_gaussian_s3_t21": lambda: gaussian_conv(3, 21, 2**20),
        "cfg4_bpx_L8": lambda: bpx.preconditioned_half_laplace(1, 8, 2),
        "cfg4_pinv_bpx_L4": lambda: ut.Pseudoinverse(bpx.preconditioned_half_laplace(1, 4, 2), 15, 0.1),
        "cfg5_bits27": lambda: cfg5(27),
        "cfg5_bits25": lambda: cfg5(25),
        "cfg5_bits26": lambda: cfg5(26),
        "cfg5_bits24": lambda: cfg5(24),
    }
    if args.columns:
        # verifier.py:134-155: besides node itself, ut.verify simulates Adjoint(node) and the controlled variant
        legs = ["laplace_n3", "laplace_n4", "laplace_readme_n4", "gaussian_s3_t4", "bpx_L2", "bpx_L3", "cfg5_bits5", "qsvt_inc2_cheb",
                "constant_vector5", "fourier4", "scale_complex", "const_int_add_4_m3", "pinv_bpx_L2"]
        for name in legs:
            if args.only and args.only != name:
                continue
            node = small[name]()
            record(name + "__adjoint", ut.Adjoint(node))
            try:  # verifier.py:146-153 (the reference only does this for nodes with their own controlled circuit)
                controlled = (node.normalization * ut.Identity(ut.Subspace("#" * node.subspace_out.total_qubits))) | node
                record(name + "__controlled", controlled)
            except Exception as exc:  # noqa: BLE001 - shapes that the reference's "|" rejects
                print(f"{name + '__controlled':34s} skipped: {type(exc).__name__}: {exc}")
        for name, make in list(mid.items()) + [("cfg5_bits13", small["cfg5_bits13"])] + list(big.items()):
            if args.only and args.only != name:
                continue
            if name.startswith("cfg5_bits2"):
                continue  # 2^30+ output entries per column: config 5 is pinned at bits=13
            record_columns(name, make())
        return
    for name, make in small.items():
        if args.only and args.only != name:
            continue
        record(name, make())
    for name, make in mid.items():
        if args.only and args.only != name:
            continue
        record(name, make(), expected=False)
    if args.big:
        for name, make in big.items():
            if args.only and args.only != name:
                continue
            record(name, make(), expected=False)


if __name__ == "__main__":
    main()

"""The reference's OWN test-suite, run through this repository's host path on CPU.

Only possible in the build container (``/root/reference`` is not shipped to the GPU box): the
reference package is imported against the tequila stand-in, ``unitaria_b200.install()`` patches
``Circuit.simulate`` / ``Node.simulate`` / ``Node.simulate_norm`` exactly as a user would, and the CUDA
engine is swapped for the numpy test double in tests/fake_engine.py.  So every ``ut.verify`` of the
reference (compute() vs simulate, atol 1e-8, verifier.py:115-126) exercises OUR lowering, subspace
compilation, projection order and adapter quirks; the kernels themselves are covered by the GPU tests.
"""

import os
import subprocess
import sys

import pytest

from tests.conftest import REFERENCE_SRC, ROOT, has_reference

pytestmark = pytest.mark.reference

PLUGIN = '''
import sys
import pytest

@pytest.fixture(scope="session", autouse=True)
def _b200_backend():
    import unitaria_b200
    from unitaria_b200 import adapter
    from tests.fake_engine import FakeStateVector
    adapter.StateVector = FakeStateVector      # test double instead of the CUDA engine
    unitaria_b200.install()
    unitaria_b200.seed_sampling(1234)
    import tequila
    calls_before = len(tequila.SIM_CALLS)
    yield
    # the reference's seam to tequila must never have been used: everything went through our path
    assert len(tequila.SIM_CALLS) == calls_before, "tq.simulate was called: the adapter was bypassed"
    unitaria_b200.uninstall()
'''


@pytest.mark.skipif(not has_reference(), reason="/root/reference exists only in the build container")
def test_reference_suite_passes_through_our_host_path(tmp_path):
    plugin_dir = tmp_path / "plug"
    plugin_dir.mkdir()
    (plugin_dir / "b200_plugin.py").write_text(PLUGIN)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(plugin_dir), ROOT, os.path.join(ROOT, "tests", "_tequila_shim"), REFERENCE_SRC])
    cmd = [
        sys.executable, "-m", "pytest", "/root/reference/tests", "/root/reference/examples", "-q", "-x",
        "-p", "no:cacheprovider", "-p", "b200_plugin", "--import-mode=importlib", "-c", "/dev/null",
        "-o", "python_files=test_*.py example_*.py",
        # estimator/test_simulator.py never touches the simulation seam (binomial model + tequila's gate
        # compiler for counting, which the stand-in does not provide); test_backend_estimator.py does
        "--ignore=/root/reference/tests/estimator/test_simulator.py",
        "--rootdir", str(tmp_path),
    ]
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=1500)
    tail = "\n".join(proc.stdout.splitlines()[-25:])
    assert proc.returncode == 0, tail + proc.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail, tail

"""Shared test plumbing.

* registers the ``gpu`` marker (``-m gpu`` tests need a B200; everything else runs on CPU);
* puts the repo root on ``sys.path`` and, when the real ``tequila`` is not importable (it is not in
  the build image nor on the GPU box), the stand-in under ``tests/_tequila_shim``;
* ``/root/reference`` exists only in the build container: tests that import the reference skip
  themselves elsewhere.
"""

import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if importlib.util.find_spec("tequila") is None:
    sys.path.insert(0, os.path.join(ROOT, "tests", "_tequila_shim"))

REFERENCE_SRC = "/root/reference/src"
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")
    config.addinivalue_line("markers", "reference: imports /root/reference (build container only)")


def has_reference() -> bool:
    return os.path.isdir(REFERENCE_SRC)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR

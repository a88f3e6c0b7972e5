"""Shared pytest configuration: marker registration, repo paths, golden-fixture loader."""
import glob
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
ORACLE_DIR = os.path.join(REPO, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)
GOLDEN = os.path.join(REPO, "tests", "golden")
REFERENCE = os.environ.get("REF_PATH", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


def load_golden(path):
    with np.load(path) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def reference_objective():
    """The live reference module (only in the build container); tests skip when absent."""
    if not os.path.isfile(os.path.join(REFERENCE, "objective.py")):
        pytest.skip("reference checkout not present on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_objective", os.path.join(REFERENCE, "objective.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

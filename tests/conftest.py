import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once (no-op when up to date; the GPU box uses the shipped .so files)."""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.load()
    return pyoracle


@pytest.fixture(scope="session")
def rp():
    import rust_pathtracer_b200
    return rust_pathtracer_b200


@pytest.fixture(scope="session")
def demo_export(rp):
    return rp.AnalyticalScene.new().device_export()


@pytest.fixture(scope="session")
def oracle_demo(po, demo_export):
    return po.OracleScene(demo_export)


@pytest.fixture(scope="session")
def oracle_literal(po):
    return po.OracleScene(None)


def unit_vectors(rng, n, dtype=np.float32):
    v = rng.normal(size=(3, n))
    v /= np.linalg.norm(v, axis=0, keepdims=True)
    return np.ascontiguousarray(v, dtype=dtype)


def rel_err(a, b, floor=1e-7):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def vec_rel_err(a, b, floor=1e-7):
    """relative error of (3, n) vectors measured against the vector norm"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b, axis=0) / np.maximum(np.linalg.norm(b, axis=0), floor)

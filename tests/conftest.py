import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def ulp_diff(a, b):
    """Element-wise distance in float32 units in the last place."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


def f32_mismatch(a, b, abs_floor=1e-9):
    """Boolean mask of elements that differ by more than one float32 ulp (of the larger magnitude) AND by more
    than ``abs_floor`` in absolute terms.  The floor matters for components of unit vectors that are ~1e-17 in
    one libm and exactly 0 in another: many 'ulps' apart, physically identical."""
    a = np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)
    b = np.ascontiguousarray(b, dtype=np.float32).astype(np.float64)
    tol = np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)).astype(np.float64), abs_floor)
    return np.abs(a - b) > tol


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def lrc():
    import lrc_b200
    return lrc_b200


@pytest.fixture(scope="session")
def golden_poses(golden):
    p = golden("poses.npz")
    return {"identity": np.eye(4), "posed": p["pose"][1].copy()}


@pytest.fixture(scope="session")
def engine(lrc):
    """A live GPU engine; only `-m gpu` tests may request it."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return lrc.RaycastEngineGPU()

import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def icp():
    """The product binding; building the library first if it is missing (nvcc cross-compiles)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_b200icp_build", os.path.join(ROOT, "3dtk_b200", "build.py"))
    build = importlib.util.module_from_spec(spec)   # by path: the package __init__ raises without the library
    spec.loader.exec_module(build)
    build.build()
    return importlib.import_module("3dtk_b200")


@pytest.fixture(scope="session")
def port():
    import orclib
    return orclib.port()


@pytest.fixture(scope="session")
def ref():
    import orclib
    L = orclib.ref()
    if L is None:
        pytest.skip("oracle/_ref/libref3dtk.so not built (needs /root/reference)")
    return L


@pytest.fixture(scope="session")
def ctx(icp):
    c = icp.Context(0)
    yield c
    c.close()


def make_pair(icp, n_model, n_data, seed=0, pos=(12.0, -7.0, 5.0), theta_deg=(0.5, -1.0, 0.8), noise=0.5):
    """Synthetic scan pair of SURVEY 8d: same geometry, independent samplings, data moved by inv(P)."""
    model = icp.synth_scene(7, 42 + seed, n_model, noise)
    data = icp.synth_scene(7, 43 + seed, n_data, noise)
    P = icp.euler_to_matrix4(np.array(pos), np.deg2rad(np.array(theta_deg)))
    Pinv, ok = icp.m4inv(P)
    assert ok == 1
    return model, icp.transform_points(Pinv, data), P

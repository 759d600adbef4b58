"""CPU-side tests of the product's host logic: the C-ABI library loads and exports every symbol the
header declares, the moment-based solvers (solve.h via b200icp_align_pairs) agree with the oracle's
pair-walking Align, the 4x4 helpers and the scene generator behave.  No GPU compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orclib
from orclib import P

ROOT = orclib.ROOT


def test_library_exports_every_declared_symbol(icp):
    hdr = open(os.path.join(ROOT, "include", "b200icp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(b200icp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(icp.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, "declared in include/b200icp.h but not exported: %s" % missing
    assert sorted(icp.EXPORTED) == declared, "python binding and header disagree"


def test_no_device_is_a_loud_error(icp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(icp.B200ICPError) as ei:
        icp.Context(0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_m4_helpers_match_oracle(icp, port):
    rng = np.random.default_rng(3)
    for _ in range(20):
        pos, th = rng.uniform(-500, 500, 3), rng.uniform(-np.pi, np.pi, 3)
        a, b = icp.euler_to_matrix4(pos, th), np.empty(16)
        port.orc_euler_to_matrix4(P(pos), P(th), P(b))
        assert np.array_equal(a, b)
        inv_a, ok = icp.m4inv(a)
        inv_b = np.empty(16)
        assert port.orc_m4inv(P(b), P(inv_b)) == 1 and ok == 1
        np.testing.assert_allclose(inv_a, inv_b, rtol=0, atol=1e-12 * (1 + np.abs(inv_b).max()))
        m2 = icp.euler_to_matrix4(rng.uniform(-5, 5, 3), rng.uniform(-1, 1, 3))
        c = np.empty(16)
        port.orc_mmult(P(a), P(m2), P(c))
        assert np.array_equal(icp.mmult(a, m2), c)
    sing = np.zeros(16)
    out, ok = icp.m4inv(sing)
    assert ok == 0 and np.array_equal(out, np.eye(4).reshape(16))   # M4inv's identity-on-failure rule


def _pairs(rng, n, theta, pos, noise, scale=500.0):
    p2 = rng.uniform(-scale, scale, (n, 3)) + np.array([3000.0, -1200.0, 800.0])  # far from the origin
    M = np.empty(16)
    orclib.port().orc_euler_to_matrix4(P(np.asarray(pos, float)), P(np.asarray(theta, float)), P(M))
    Mm = M.reshape(4, 4).T
    p1 = p2 @ Mm[:3, :3].T + Mm[:3, 3] + rng.normal(0, noise, (n, 3))
    nr = rng.normal(size=(n, 3))
    nr /= np.linalg.norm(nr, axis=1, keepdims=True)
    return np.ascontiguousarray(p1), np.ascontiguousarray(p2), np.ascontiguousarray(nr)


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5, 6, 10])
@pytest.mark.parametrize("n", [4, 50, 20000])
def test_align_pairs_matches_oracle_align(icp, port, algo, n):
    rng = np.random.default_rng(100 * algo + n)
    p1, p2, nr = _pairs(rng, n, [0.01, -0.02, 0.015], [3.0, -2.0, 1.0], 0.4)
    cm, cdv = p1.mean(0), p2.mean(0)
    want = np.zeros(16)
    r_want = port.orc_align(algo, n, P(p1), P(p2), P(nr), P(cm), P(cdv), 0, P(want))
    got, r_got = icp.align_pairs(algo, p1, p2, nr if algo == 10 else None, cm, cdv)
    if r_want == -1.0:
        assert r_got == -1.0
        return
    assert abs(r_got - r_want) <= 1e-12 * max(1.0, abs(r_want))
    # tolerance: 1e-9 relative Frobenius (the north-star gate on whole matches is 1e-4); ORTHO forms H^T H, which
    # squares the condition number -- with 4 pairs its smallest eigenvalue amplifies the rounding difference
    # between centred pair sums (reference / oracle) and moments (product)
    assert orclib.rel_frobenius(got, want) < (1e-6 if algo == 3 and n < 10 else 1e-9)


def test_align_pairs_reflection_and_degenerate(icp, port):
    # planar, noisy input: exercises the det(R) < 0 repair of icp6D_SVD (icp6Dsvd.cc:101-115)
    rng = np.random.default_rng(5)
    p2 = np.c_[rng.uniform(-100, 100, (200, 2)), np.zeros(200)]
    p1 = p2 + rng.normal(0, 5.0, p2.shape)
    p1[:, 2] = 0.0
    cm, cdv = p1.mean(0), p2.mean(0)
    for algo in (1, 2):
        want = np.zeros(16)
        port.orc_align(algo, 200, P(p1), P(p2), None, P(cm), P(cdv), 0, P(want))
        got, _ = icp.align_pairs(algo, p1, p2, None, cm, cdv)
        R = got.reshape(4, 4).T[:3, :3]
        assert abs(np.linalg.det(R) - 1.0) < 1e-9
        assert orclib.rel_frobenius(got, want) < 1e-7
    # APX with <= 3 pairs: identity, rms 0 (icp6Dapx.cc:42-46)
    got, r = icp.align_pairs(6, p1[:3], p2[:3])
    assert r == 0.0 and np.array_equal(got, np.eye(4).reshape(16))
    # Cholesky failure is reported as -1 (icp6Dapx.cc:97-100): all data points identical -> A == 0
    same = np.tile(p2[:1], (10, 1))
    got, r = icp.align_pairs(6, same + 1.0, same)
    assert r == -1.0


def test_synth_scene_is_deterministic_and_bounded(icp):
    a = icp.synth_scene(7, 42, 5000, 0.5)
    b = icp.synth_scene(7, 42, 5000, 0.5)
    c = icp.synth_scene(7, 43, 5000, 0.5)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a[:, 0].min() > -1010 and a[:, 0].max() < 1010
    assert a[:, 1].min() > -10 and a[:, 1].max() < 310
    assert a[:, 2].min() > -510 and a[:, 2].max() < 510
    clean = icp.synth_scene(7, 42, 5000, 0.0)
    on_shell = (np.abs(np.abs(clean[:, 0]) - 1000) < 1e-9) | (np.abs(clean[:, 1]) < 1e-9) | \
               (np.abs(clean[:, 1] - 300) < 1e-9) | (np.abs(np.abs(clean[:, 2]) - 500) < 1e-9)
    assert 0.5 < on_shell.mean() < 1.0   # most samples on the room shell, the rest on walls / boxes

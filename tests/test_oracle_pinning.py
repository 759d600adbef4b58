"""Pins the oracle (oracle/oracle_icp.cpp, our CPU restatement) before anything trusts it:
  1. the reference's own known-answer tests for the k-d tree (testing/kdtree/kdtree.cc:20-46) and its
     seeded differential pattern (testing/kdtree/kdtree_indexed_random.cc:14-26, :192-220);
  2. golden vectors generated from the UNMODIFIED reference objects (tests/golden/make_golden.py ->
     tests/golden/ref_vectors.npz): NN indices, getPtPairs pair lists, the four Align results, whole
     matches, k-NN normals;
  3. live differential against oracle/_ref/libref3dtk.so when it is present.
CPU only."""
import os

import numpy as np
import pytest

import orclib
from orclib import P

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))


def _tree(port, pts):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    return pts, port.orc_tree_create(P(pts), len(pts), 20)


# ---- 1. reference KATs ------------------------------------------------------------------------
def test_kat_exactly_maxdist2_is_rejected(port):       # kdtree.cc:20-26
    pts, t = _tree(port, [[2.0, 0.0, 0.0]])
    assert port.orc_find_closest(t, P(np.zeros(3)), 4.0) == -1


def test_kat_just_inside_is_accepted(port):            # kdtree.cc:29-35
    pts, t = _tree(port, [[1.99999999999, 0.0, 0.0]])
    assert port.orc_find_closest(t, P(np.zeros(3)), 4.0) == 0


def test_kat_nearer_of_two(port):                      # kdtree.cc:39-46
    pts, t = _tree(port, [[1.5, 0.0, 0.0], [1.0, 0.0, 0.0]])
    assert port.orc_find_closest(t, P(np.zeros(3)), 4.0) == 1


def test_seeded_differential_tree_vs_bruteforce(port):  # kdtree_indexed_random.cc:192-220
    rng = np.random.default_rng(42)
    pts, t = _tree(port, rng.uniform(-10, 10, (10000, 3)))
    for md2 in np.arange(0.5, 5.01, 0.5):
        for _ in range(100):
            q = rng.uniform(-10, 10, 3)
            assert port.orc_find_closest(t, P(q), float(md2)) == port.orc_brute_closest(P(pts), len(pts), P(q), float(md2))


def test_zero_points_is_an_error(port):                 # kdTreeImpl.h:86-88
    assert port.orc_tree_create(None, 0, 20) is None


# ---- 2. golden vectors from the compiled reference --------------------------------------------
def test_golden_nn_indices(port):
    pts, t = _tree(port, GOLD["nn_points"])
    q = np.ascontiguousarray(GOLD["nn_queries"])
    for r, md2 in enumerate(GOLD["nn_maxdist2"]):
        idx = np.empty(len(q), np.int32)
        port.orc_find_closest_batch(t, P(q), len(q), float(md2), P(idx), None)
        assert np.array_equal(idx, GOLD["nn_idx_KDtree_FindClosest"][r])


@pytest.mark.parametrize("mode", [0, 2])
def test_golden_get_pt_pairs(mode):
    tree = orclib.PortTree(GOLD["pair_model"])
    S = np.ascontiguousarray(GOLD["pair_source_alignxf"])
    data, nrm = np.ascontiguousarray(GOLD["pair_data"]), np.ascontiguousarray(GOLD["pair_data_normals"])
    k, p1, p2, pn, idx, sm, cm, cd = tree.get_pt_pairs(S, data, nrm, 400.0, mode)
    assert k == len(GOLD["pairs%d_p1" % mode])
    assert np.array_equal(p1, GOLD["pairs%d_p1" % mode])       # bit-exact: same arithmetic, same order
    assert np.array_equal(p2, GOLD["pairs%d_p2" % mode])
    if mode:
        assert np.array_equal(pn, GOLD["pairs%d_n" % mode])
    assert np.array_equal(np.r_[sm, cm, cd], GOLD["pairs%d_sum_cm_cd" % mode])


@pytest.mark.parametrize("mode,algo", [(0, 1), (0, 2), (0, 3), (0, 4), (0, 5), (0, 6), (2, 1), (2, 10)])
def test_golden_align(port, mode, algo):
    p1 = np.ascontiguousarray(GOLD["pairs%d_p1" % mode]); p2 = np.ascontiguousarray(GOLD["pairs%d_p2" % mode])
    pn = np.ascontiguousarray(GOLD["pairs%d_n" % mode])
    s = GOLD["pairs%d_sum_cm_cd" % mode]
    k = len(p1)
    cm, cd = np.ascontiguousarray(s[1:4] / k), np.ascontiguousarray(s[4:7] / k)
    xf = np.zeros(16)
    rms = port.orc_align(algo, k, P(p1), P(p2), P(pn), P(cm), P(cd), 0, P(xf))
    want = GOLD["align_mode%d_algo%d" % (mode, algo)]
    assert abs(rms - want[16]) <= 1e-13 * abs(want[16])
    assert orclib.rel_frobenius(xf, want[:16]) < 1e-12


@pytest.mark.parametrize("algo,mode", [(1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (10, 2), (1, 2)])
def test_golden_match(algo, mode):
    md, it, eps = GOLD["match_maxdist_iters_eps"]
    r = orclib.port_match(GOLD["pair_model"], GOLD["pair_data"], GOLD["pair_data_normals"] if mode else None,
                          algo=algo, mode=mode, max_dist=float(md), max_iter=int(it), eps=float(eps))
    key = "match_algo%d_mode%d_" % (algo, mode)
    assert r["iterations"] == int(GOLD[key + "iterations"][0])
    assert np.array_equal(r["npairs"], GOLD[key + "npairs"])
    np.testing.assert_allclose(r["rms"], GOLD[key + "rms"], rtol=1e-10)
    assert orclib.rel_frobenius(r["transmat"], GOLD[key + "transmat"]) < 1e-10


def test_golden_normals(port):
    pts = np.ascontiguousarray(GOLD["normals_points"]); rpos = np.ascontiguousarray(GOLD["normals_rpos"])
    out = np.empty_like(pts)
    port.orc_normals_knn(P(pts), len(pts), 10, P(rpos), P(out))
    dots = (out * GOLD["normals_calculateNormalsKNN_k10"]).sum(1)
    # eigenvectors of (nearly) degenerate neighbourhoods are not unique; everything else must agree
    assert (dots > 1 - 1e-9).mean() > 0.995 and (np.abs(dots) > 1 - 1e-6).mean() > 0.999


# ---- 3. live differential against the compiled reference --------------------------------------
def test_live_nn_incl_tie_order(port, ref):
    g = np.stack(np.meshgrid(*[np.arange(10.0)] * 3, indexing="ij"), -1).reshape(-1, 3).copy()
    q = np.ascontiguousarray(g[:400] + 0.5)
    _, t = _tree(port, g)
    rt = ref.ref_tree_create(P(g), len(g), 0, 20)
    a, b = np.empty(len(q), np.int32), np.empty(len(q), np.int32)
    port.orc_find_closest_batch(t, P(q), len(q), 4.0, P(a), None)
    ref.ref_find_closest_batch(rt, P(q), len(q), 4.0, P(b), 1)
    assert np.array_equal(a, b)       # same build / traversal order -> same winner among 8 ties
    ref.ref_tree_free(rt)


def test_live_math_helpers(port, ref):
    rng = np.random.default_rng(9)
    for _ in range(50):
        pos, th = rng.uniform(-1000, 1000, 3), rng.uniform(-3, 3, 3)
        a, b = np.empty(16), np.empty(16)
        port.orc_euler_to_matrix4(P(pos), P(th), P(a)); ref.ref_euler_to_matrix4(P(pos), P(th), P(b))
        assert np.array_equal(a, b)
        ia, ib = np.empty(16), np.empty(16)
        assert port.orc_m4inv(P(a), P(ia)) == ref.ref_m4inv(P(b), P(ib)) == 1
        assert np.array_equal(ia, ib)
        c, d = np.empty(16), np.empty(16)
        port.orc_mmult(P(a), P(ia), P(c)); ref.ref_mmult(P(b), P(ib), P(d))
        assert np.array_equal(c, d)


@pytest.mark.parametrize("algo,mode", [(1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0), (10, 2)])
def test_live_match_vs_reference(ref, algo, mode):
    rng = np.random.default_rng(77 + algo)
    base = rng.uniform(-200, 200, (6000, 3)); base[:, 1] = np.abs(base[:, 1]) * 0.2
    model = base + rng.normal(0, 0.3, base.shape)
    th = np.deg2rad([0.6, -0.9, 0.7]); c, s = np.cos(th[2]), np.sin(th[2])
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    data = np.ascontiguousarray((base[:5000] + rng.normal(0, 0.3, (5000, 3))) @ Rz.T + np.array([3.0, 1.0, -2.0]))
    nrm = np.ascontiguousarray(rng.normal(size=data.shape)) if mode else None
    a = orclib.port_match(model, data, nrm, algo=algo, mode=mode, max_dist=15.0, max_iter=25, eps=1e-6)
    b = orclib.ref_match(model, data, nrm, algo=algo, mode=mode, max_dist=15.0, max_iter=25, eps=1e-6)
    assert a["iterations"] == b["iterations"] and np.array_equal(a["npairs"], b["npairs"])
    assert orclib.rel_frobenius(a["transmat"], b["transmat"]) < 1e-10
    np.testing.assert_allclose(a["xyz"], b["xyz"], rtol=0, atol=1e-9)


def test_live_parallel_arm_reaches_the_same_fixed_point(ref):
    """The OpenMP arm (Align_Parallel, icp6Dquat.cc:515-634) is what the CPU baseline times; its
    per-iteration transforms differ slightly from the serial arm but the converged pose must agree."""
    L = orclib.ref(omp=True)
    if L is None:
        pytest.skip("libref3dtk_omp.so not built")
    rng = np.random.default_rng(3)
    base = rng.uniform(-200, 200, (8000, 3)); base[:, 1] = np.abs(base[:, 1]) * 0.2
    model = base + rng.normal(0, 0.3, base.shape)
    data = np.ascontiguousarray(base[:7000] + rng.normal(0, 0.3, (7000, 3)) + np.array([2.0, -1.0, 1.5]))
    ser = orclib.ref_match(model, data, algo=1, max_dist=15.0, max_iter=60, eps=1e-7)
    par = orclib.ref_match(model, data, algo=1, max_dist=15.0, max_iter=60, eps=1e-7, threads=4, omp=True)
    assert orclib.rel_frobenius(par["transmat"], ser["transmat"]) < 1e-4


def test_live_lum_link_vs_harness_with_reference_newmat(port, ref):
    """covarianceEuler: our Gaussian elimination vs the reference's newmat inverse (harness restatement;
    lum6Deuler.cc itself does not compile here -- see oracle/oracle_icp.cpp)."""
    rng = np.random.default_rng(12)
    base = rng.uniform(-300, 300, (8000, 3)); base[:, 1] = np.abs(base[:, 1]) * 0.3
    model = np.ascontiguousarray(base + rng.normal(0, 0.3, base.shape))
    data = np.ascontiguousarray(base[:6000] + rng.normal(0, 0.3, (6000, 3)) + [1.0, -0.5, 0.7])
    S = np.empty(16)
    port.orc_euler_to_matrix4(P(np.array([0.4, 0.2, -0.3])), P(np.deg2rad([0.1, -0.2, 0.15])), P(S))
    C1, D1, m1 = orclib.port_lum_link(model, data, 100.0, S)
    rt = ref.ref_tree_create(P(model), len(model), 0, 20)
    C2, D2 = np.zeros(36), np.zeros(6)
    m2 = ref.ref_lum_link(rt, P(S), P(data), len(data), 100.0, P(C2), P(D2))
    ref.ref_tree_free(rt)
    assert m1 == m2 and m1 > 1000
    np.testing.assert_allclose(C1.reshape(-1), C2, rtol=1e-9)
    np.testing.assert_allclose(D1, D2, rtol=1e-9, atol=1e-9 * np.abs(D2).max())
    # identical clouds -> zero information (lum6Deuler.cc:219-231)
    C3, D3, m3 = orclib.port_lum_link(model, model, 100.0)
    assert m3 == len(model) and not C3.any() and not D3.any()


def test_parallel_search_serial_align_arm_is_bit_identical_to_the_serial_reference():
    """ref_match(parallel_threads < 0) -- the full-size parity oracle of bench.py and tests/test_gpu_parity.py --
    must leave exactly what the serial arm (icp6D.cc:224-244) leaves: same pairs in the same order, same sums."""
    import importlib
    if orclib.ref(omp=True) is None:
        pytest.skip("oracle/_ref/libref3dtk_omp.so not built (needs /root/reference)")
    icp = importlib.import_module("3dtk_b200")
    from conftest import make_pair
    model, data, _ = make_pair(icp, 30000, 25001)
    a = orclib.ref_match(model, data, algo=1, threads=0, omp=False)     # serial library, serial arm
    for th in (-3, -8):
        b = orclib.ref_match(model, data, algo=1, threads=th, omp=True)
        assert a["iterations"] == b["iterations"]
        assert np.array_equal(a["npairs"], b["npairs"]) and np.array_equal(a["rms"], b["rms"])
        assert np.array_equal(a["transmat"], b["transmat"]) and np.array_equal(a["xyz"], b["xyz"])

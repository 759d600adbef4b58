"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle on the same inputs.

Bars: NN indices / found flags / pair counts bit-exact (index work); distances bit-exact in exact mode;
transforms of whole matches within 1e-4 relative Frobenius (north_star), in practice ~1e-9.
"""
import os

import numpy as np
import pytest

import orclib
from orclib import P
from conftest import make_pair

pytestmark = pytest.mark.gpu

TOL_MATCH = 1e-4   # north_star: transforms within 1e-4 rel Frobenius of reference icp6D_QUAT


# ------------------------------------------------------------------ reference KATs on the GPU tree
# testing/kdtree/kdtree.cc:20-46
def test_kat_radius_boundary_is_strict(icp, ctx):
    s = icp.Scan(ctx, np.array([[2.0, 0.0, 0.0]]))
    assert s.find_closest([0.0, 0.0, 0.0], 4.0) == -1            # exactly maxdist2 away: rejected
    s = icp.Scan(ctx, np.array([[1.99999999999, 0.0, 0.0]]))
    assert s.find_closest([0.0, 0.0, 0.0], 4.0) == 0             # just inside: accepted
    s = icp.Scan(ctx, np.array([[1.5, 0.0, 0.0], [1.0, 0.0, 0.0]]))
    assert s.find_closest([0.0, 0.0, 0.0], 4.0) == 1             # nearer of two


# testing/kdtree/kdtree_indexed_random.cc:192-220 -- same shape of test: 10 000 uniform points in
# [-10,10]^3, 10 radii x 100 queries, result must equal the brute-force oracle (index equality)
def test_seeded_differential_vs_bruteforce(icp, ctx, port):
    rng = np.random.default_rng(42)
    pts = rng.uniform(-10, 10, (10000, 3))
    s = icp.Scan(ctx, pts)
    for md2 in np.arange(0.5, 5.01, 0.5):
        q = rng.uniform(-10, 10, (100, 3))
        idx, d2, _ = s.nn_batch(q, float(md2))
        for i in range(len(q)):
            want = port.orc_brute_closest(P(pts), len(pts), P(q[i]), float(md2))
            assert idx[i] == want
            if want >= 0:
                assert d2[i] == ((pts[want] - q[i]) ** 2).sum() or abs(d2[i] - ((pts[want] - q[i]) ** 2).sum()) < 1e-15


@pytest.mark.parametrize("cell_edge", [0.0, 0.7, 3.0, 40.0])
@pytest.mark.parametrize("maxdist", [0.3, 2.0, 25.0, 1.0e4])
def test_nn_batch_equals_kdtree_oracle(icp, ctx, cell_edge, maxdist):
    """Stage 1 only (big cells), stage 2 dominated (tiny cells), unbounded radius, auto cell edge."""
    rng = np.random.default_rng(7)
    model = icp.synth_scene(7, 1, 30000, 0.5)
    q = np.r_[icp.synth_scene(7, 2, 3000, 0.5) + rng.normal(0, 3.0, (3000, 3)),
              rng.uniform(-1500, 1500, (500, 3))]                 # some far outside the bbox
    s = icp.Scan(ctx, model, cell_edge=cell_edge, max_dist_hint=maxdist)
    idx, d2, sums = s.nn_batch(q, maxdist * maxdist)
    tree = orclib.PortTree(model)
    want_idx, want_d2 = tree.find_closest_batch(q, maxdist * maxdist)
    assert np.array_equal(idx >= 0, want_idx >= 0)
    found = want_idx >= 0
    assert np.array_equal(d2[found], want_d2[found])              # bit-exact fp64 distances
    same = idx[found] == want_idx[found]
    # different index only allowed on exact distance ties
    assert same.all() or np.array_equal(d2[found][~same], want_d2[found][~same])
    assert sums[0] == found.sum()


def test_ties_resolve_to_lowest_row(icp, ctx):
    g = np.stack(np.meshgrid(*[np.arange(8.0)] * 3, indexing="ij"), -1).reshape(-1, 3).copy()
    s = icp.Scan(ctx, g, cell_edge=1.5)
    q = g[:200] + 0.5          # 8 equidistant corners each
    idx, d2, _ = s.nn_batch(q, 4.0)
    d2_all = ((q[:, None, :] - g[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(idx, d2_all.argmin(1))                  # argmin = first (lowest row) minimum
    assert np.allclose(d2, 0.75)


def test_get_pt_pairs_sums_and_plane_projection(icp, ctx, ref):
    """SearchTree::getPtPairs with a moved model (source_alignxf != I) against the compiled reference."""
    rng = np.random.default_rng(11)
    model = icp.synth_scene(7, 5, 20000, 0.5)
    data = icp.synth_scene(7, 6, 4000, 0.5)
    nrm = rng.normal(size=data.shape) * 3.0                       # un-normalised on purpose
    S = icp.euler_to_matrix4(np.array([4.0, -3.0, 2.0]), np.deg2rad([0.7, -0.4, 1.1]))
    data_g = icp.transform_points(S, data)                        # data roughly in the moved model frame
    s = icp.Scan(ctx, model)
    rt = ref.ref_tree_create(P(model), len(model), 0, 20)
    for mode in (0, 2):
        idx, d2, sums = s.nn_batch(data_g, 625.0, source_alignxf=S, q_nrm=nrm if mode else None,
                                   pairing_mode=mode)
        n = len(data_g)
        p1, p2, pn = np.empty((n, 3)), np.empty((n, 3)), np.empty((n, 3))
        sm, cm, cdv = np.zeros(1), np.zeros(3), np.zeros(3)
        k = ref.ref_get_pt_pairs(rt, P(S), P(data_g), P(nrm), 0, n, 0, 1, 625.0, mode, P(p1), P(p2), P(pn),
                                 P(sm), P(cm), P(cdv))
        assert sums[0] == k
        np.testing.assert_allclose(sums[1], sm[0], rtol=1e-11)
        np.testing.assert_allclose(sums[2:5], cm, rtol=1e-11, atol=1e-7)
        np.testing.assert_allclose(sums[5:8], cdv, rtol=1e-11, atol=1e-7)
        # pairs themselves: model rows picked == rows the reference picked
        Sm = S.reshape(4, 4).T
        mine = model[idx[idx >= 0]] @ Sm[:3, :3].T + Sm[:3, 3]
        if mode == 0:
            np.testing.assert_allclose(mine, p1[:k], rtol=0, atol=1e-9)
    ref.ref_tree_free(rt)


# ------------------------------------------------------------------ fused match vs oracle match
@pytest.fixture(params=["fused", "split"])
def iteration_form(request, monkeypatch):
    """Both device forms of one ICP iteration: the single fused kernel (default) and the two-kernel form
    (TMA-streamed pass + queued searches, stream_kernels.cuh), selected per call by B200ICP_SPLIT."""
    monkeypatch.setenv("B200ICP_SPLIT", "1" if request.param == "split" else "0")
    return request.param


@pytest.mark.parametrize("algo", [1, 2, 3, 4, 5, 6])
def test_match_point_to_point_vs_oracle(icp, ctx, algo, iteration_form):
    model, data, Ptrue = make_pair(icp, 60000, 50000)
    want = orclib.port_match(model, data, algo=algo, max_dist=25.0, max_iter=50, eps=1e-5)
    m, d = icp.Scan(ctx, model, max_dist_hint=25.0), icp.Scan(ctx, data, max_dist_hint=25.0)
    got = icp.icp6D(ctx, algo=algo, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5).match(m, d)
    T, D = d.get_pose()
    assert got["iterations"] == want["iterations"]
    assert np.array_equal(got["npairs"], want["npairs"])          # same pair sets every iteration
    np.testing.assert_allclose(got["rms"], want["rms"], rtol=1e-9)
    assert orclib.rel_frobenius(T, want["transmat"]) < 1e-8 < TOL_MATCH
    assert orclib.rel_frobenius(D, want["dalignxf"]) < 1e-8
    assert orclib.rel_frobenius(T, Ptrue) < 2e-2                  # and it actually registered the pair
    moved = d.download()
    np.testing.assert_allclose(moved, want["xyz"], rtol=0, atol=1e-8)


@pytest.mark.parametrize("scale,offset", [(0.01, 0.0), (1000.0, 0.0), (1.0, 5.0e4)])
def test_match_in_other_units_and_far_from_the_origin(icp, ctx, scale, offset):
    """The point-to-point kernel sums its pair moments as 64-bit integers with per-match power-of-two scales taken from
    the scene's extent (order-independent sums, icp_kernels.cuh): metres instead of centimetres, a 1000x larger scene
    and a scene 500 m away from the origin must register like the oracle, pair for pair."""
    model, data, _ = make_pair(icp, 30000, 30000)
    model, data = model * scale + offset, data * scale + offset
    md, eps = 25.0 * scale, 1e-5 * scale
    want = orclib.port_match(model, data, algo=1, max_dist=md, max_iter=30, eps=eps)
    m, d = icp.Scan(ctx, model, max_dist_hint=md), icp.Scan(ctx, data, max_dist_hint=md)
    got = icp.icp6D(ctx, algo=1, max_dist_match=md, max_num_iterations=30, epsilon_icp=eps).match(m, d)
    T, _ = d.get_pose()
    assert got["iterations"] == want["iterations"]
    assert np.array_equal(got["npairs"], want["npairs"])
    np.testing.assert_allclose(got["rms"], want["rms"], rtol=1e-8)
    assert orclib.rel_frobenius(T, want["transmat"]) < 1e-8 < TOL_MATCH


def test_match_fast_mode_within_north_star_tolerance(icp, ctx):
    model, data, _ = make_pair(icp, 60000, 50000)
    want = orclib.port_match(model, data, algo=1, max_dist=25.0, max_iter=50, eps=1e-5)
    m, d = icp.Scan(ctx, model, max_dist_hint=25.0), icp.Scan(ctx, data, max_dist_hint=25.0)
    icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5, exact=False).match(m, d)
    T, _ = d.get_pose()
    assert orclib.rel_frobenius(T, want["transmat"]) < TOL_MATCH


@pytest.mark.parametrize("algo", [10, 1])
def test_match_point_to_plane_vs_oracle(icp, ctx, algo, iteration_form):
    model, data, Ptrue = make_pair(icp, 40000, 30000, theta_deg=(0.3, -0.5, 0.4), pos=(6.0, -3.0, 2.0))
    nrm = icp.normals_knn(ctx, data, 10, np.array([0.0, 150.0, 0.0]))
    want = orclib.port_match(model, data, nrm, algo=algo, mode=2, max_dist=25.0, max_iter=30, eps=1e-5)
    m = icp.Scan(ctx, model, max_dist_hint=25.0)
    d = icp.Scan(ctx, data, normals=nrm, max_dist_hint=25.0)
    got = icp.icp6D(ctx, algo=algo, max_dist_match=25.0, max_num_iterations=30, epsilon_icp=1e-5).match(
        m, d, icp.CLOSEST_PLANE_SIMPLE)
    T, _ = d.get_pose()
    assert got["iterations"] == want["iterations"]
    assert np.array_equal(got["npairs"], want["npairs"])
    np.testing.assert_allclose(got["rms"], want["rms"], rtol=1e-8)
    assert orclib.rel_frobenius(T, want["transmat"]) < 1e-7 < TOL_MATCH
    xyz, n2 = d.download(with_normals=True)
    np.testing.assert_allclose(xyz, want["xyz"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(n2, want["nrm"], rtol=0, atol=1e-9)  # normals follow transform3normal


def test_match_with_moved_model_and_sequential_roles(icp, ctx, iteration_form):
    """scan1 is matched to scan0 and then serves as the model for scan2 (its tree stays in the original
    frame, queries go through inv(dalignxf) -- 'fast corresponding points', scan.cc:1208-1240)."""
    a = icp.synth_scene(7, 10, 30000, 0.5)
    P1 = icp.euler_to_matrix4(np.array([8.0, 2.0, -4.0]), np.deg2rad([0.4, 0.6, -0.5]))
    P2 = icp.euler_to_matrix4(np.array([-5.0, 3.0, 6.0]), np.deg2rad([-0.3, 0.5, 0.6]))
    b = icp.transform_points(icp.m4inv(P1)[0], icp.synth_scene(7, 11, 30000, 0.5))
    c = icp.transform_points(icp.m4inv(P2)[0], icp.synth_scene(7, 12, 30000, 0.5))
    s0, s1, s2 = icp.Scan(ctx, a), icp.Scan(ctx, b), icp.Scan(ctx, c)
    eng = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=40, epsilon_icp=1e-5)
    eng.match(s0, s1)
    eng.match(s1, s2)
    # oracle: same chain with explicit dalignxf bookkeeping
    w1 = orclib.port_match(a, b, algo=1, max_iter=40)
    w2 = orclib.port_match(b, c, algo=1, max_iter=40, model_dalignxf=w1["dalignxf"])
    assert orclib.rel_frobenius(s1.get_pose()[0], w1["transmat"]) < 1e-8
    assert orclib.rel_frobenius(s2.get_pose()[0], w2["transmat"]) < 1e-7


# ------------------------------------------------------------------ edge cases
def test_edge_cases(icp, ctx):
    with pytest.raises(icp.B200ICPError) as ei:
        icp.Scan(ctx, np.zeros((0, 3)))
    assert ei.value.code == -5                                     # kdTreeImpl.h:86-88: zero points
    one = icp.Scan(ctx, np.array([[1.0, 2.0, 3.0]]))
    assert one.find_closest([1.0, 2.0, 3.0], 1e-30) == 0           # distance 0 < any positive radius
    assert one.find_closest([1.0, 2.0, 3.0], 0.0) == -1            # 0 < 0 is false
    dup = icp.Scan(ctx, np.tile(np.array([[5.0, 5.0, 5.0]]), (1000, 1)))   # all points identical
    idx, d2, _ = dup.nn_batch(np.array([[5.0, 5.0, 6.0], [50.0, 5.0, 5.0]]), 4.0)
    assert idx[0] == 0 and d2[0] == 1.0 and idx[1] == -1
    idx, _, sums = dup.nn_batch(np.zeros((0, 3)), 4.0)             # empty query batch
    assert len(idx) == 0 and sums[0] == 0
    nanq = np.array([[np.nan, 0.0, 0.0], [5.0, 5.0, 5.5]])
    idx, _, _ = dup.nn_batch(nanq, 4.0)
    assert idx[0] == -1 and idx[1] == 0
    # ragged sizes (not a multiple of the 256-thread tile) and fewer than 4 pairs -> loop breaks at once
    far_model, far_data = icp.Scan(ctx, np.random.default_rng(0).uniform(0, 10, (777, 3))), \
        icp.Scan(ctx, np.random.default_rng(1).uniform(1000, 1010, (333, 3)))
    r = icp.icp6D(ctx, max_dist_match=5.0, max_num_iterations=10).match(far_model, far_data)
    assert r["iterations"] == 0 and r["iterations_run"] == 0
    assert np.array_equal(far_data.get_pose()[0], np.eye(4).reshape(16))
    r = icp.icp6D(ctx, max_dist_match=5.0, max_num_iterations=0).match(far_model, far_data)
    assert r["iterations"] == 0
    with pytest.raises(ValueError):
        icp.icp6D(ctx, max_dist_match=-1.0)


def test_match_is_deterministic(icp, ctx, iteration_form):
    model, data, _ = make_pair(icp, 40000, 40000)
    outs = []
    for _ in range(2):
        m, d = icp.Scan(ctx, model), icp.Scan(ctx, data)
        r = icp.icp6D(ctx, algo=1, max_num_iterations=30, epsilon_icp=1e-5).match(m, d)
        outs.append((d.get_pose()[0].copy(), r["rms"].copy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


# ------------------------------------------------------------------ full-size properties (1M points)
def test_full_size_properties(icp, ctx, iteration_form):
    n = 1_000_000
    model, data, Ptrue = make_pair(icp, n, n)
    m, d = icp.Scan(ctx, model, max_dist_hint=25.0), icp.Scan(ctx, data, max_dist_hint=25.0)
    r = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5).match(m, d)
    T, _ = d.get_pose()
    assert r["iterations_run"] >= 3
    assert orclib.rel_frobenius(T, Ptrue) < 2e-3                   # recovers the known SE(3) offset
    assert r["rms"][-1] < r["rms"][0]
    # idempotence: a scan matched against itself does not move, every point pairs with itself
    m2 = icp.Scan(ctx, model, max_dist_hint=25.0)
    r2 = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=5, epsilon_icp=1e-5).match(m, m2)
    assert int(r2["npairs"][0]) == n and r2["rms"][0] == 0.0
    assert orclib.rel_frobenius(m2.get_pose()[0], np.eye(4).reshape(16)) < 1e-12
    # NN property at full size on a query sample: agrees with the k-d tree oracle
    q = data[:: n // 2000][:2000]
    idx, d2, _ = m.nn_batch(q, 625.0)
    tree = orclib.PortTree(model)
    wi, wd = tree.find_closest_batch(q, 625.0)
    assert np.array_equal(idx >= 0, wi >= 0) and np.array_equal(d2[wi >= 0], wd[wi >= 0])


def test_bench_config_parity_vs_compiled_reference(icp, ctx):
    """BASELINE.md section 3 gate at the bench size (configs[1], 1M x 1M, icp6D_QUAT): the fused match against the
    compiled, unmodified reference (oracle/_ref: kd.cc / searchTree.cc / icp6Dquat.cc) run on the identical arrays
    with the serial arm's arithmetic (icp6D.cc:224-244), its k-d tree searches spread over the host cores
    (oracle/ref_harness.cc, bit-identical to the serial arm: tests/test_oracle_pinning.py)."""
    if orclib.ref(omp=True) is None:
        pytest.skip("oracle/_ref/libref3dtk_omp.so not built (needs /root/reference)")
    n = 1_000_000
    model, data, _ = make_pair(icp, n, n)
    try:
        threads = len(os.sched_getaffinity(0))
    except Exception:
        threads = os.cpu_count() or 1
    want = orclib.ref_match(model, data, algo=1, max_dist=25.0, max_iter=50, eps=1e-5, threads=-min(threads, 256),
                            omp=True)
    m, d = icp.Scan(ctx, model, max_dist_hint=25.0), icp.Scan(ctx, data, max_dist_hint=25.0)
    got = icp.icp6D(ctx, algo=1, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5).match(m, d)
    T, _ = d.get_pose()
    assert got["iterations"] == want["iterations"]
    assert np.array_equal(got["npairs"], want["npairs"])           # same pair count in every iteration
    np.testing.assert_allclose(got["rms"], want["rms"], rtol=1e-9)
    err = orclib.rel_frobenius(T, want["transmat"])
    assert err < 1e-8 < TOL_MATCH, err                             # north-star gate: 1e-4


def test_normals_knn_vs_oracle(icp, ctx, port):
    pts = icp.synth_scene(7, 21, 20000, 0.5)
    rpos = np.array([0.0, 150.0, 0.0])
    got = icp.normals_knn(ctx, pts, 10, rpos)
    want = np.empty_like(pts)
    port.orc_normals_knn(P(pts), len(pts), 10, P(rpos), P(want))
    dots = np.abs((got * want).sum(1))
    assert np.allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-12)
    # identical neighbour sets -> identical covariance; eigenvectors agree to rounding except where
    # the two smallest eigenvalues are (nearly) degenerate
    assert (dots > 1 - 1e-9).mean() > 0.999
    assert ((got * (pts - rpos)).sum(1) >= -1e-9).all()            # orientation rule normals.cc:95-104


def test_lum_link_vs_oracle(icp, ctx):
    """lum6DEuler::covarianceEuler on the device (two kernel passes) against the oracle."""
    a = icp.synth_scene(7, 31, 40000, 0.5)
    Pm = icp.euler_to_matrix4(np.array([1.5, -0.8, 0.6]), np.deg2rad([0.05, -0.08, 0.06]))
    b = icp.transform_points(icp.m4inv(Pm)[0], icp.synth_scene(7, 32, 30000, 0.5))
    S = icp.euler_to_matrix4(np.array([0.3, 0.1, -0.2]), np.deg2rad([0.02, 0.01, -0.03]))
    first, second = icp.Scan(ctx, a), icp.Scan(ctx, b)
    first.set_pose(S, S)
    Cg, CDg, n = icp.lum_link(ctx, first, second, 100.0)
    Cw, CDw, m = orclib.port_lum_link(a, b, 100.0, S)
    assert n == m and m > 1000
    np.testing.assert_allclose(Cg, Cw, rtol=1e-9)
    np.testing.assert_allclose(CDg, CDw, rtol=1e-8, atol=1e-9 * np.abs(CDw).max())
    Cz, CDz, nz = icp.lum_link(ctx, first, icp.Scan(ctx, a), 100.0)   # identical clouds, moved model
    assert nz > 0


def test_scan_calc_normals_on_device(icp, ctx):
    """Scan::calcNormals on a resident scan == the host-pointer path, and a point-to-plane match fed by it == one fed
    by uploaded normals (no host round trip, no second grid build)."""
    model, data, _ = make_pair(icp, 40000, 30000, theta_deg=(0.3, -0.5, 0.4), pos=(6.0, -3.0, 2.0))
    rpos = np.array([0.0, 150.0, 0.0])
    want = icp.normals_knn(ctx, data, 10, rpos)
    d1 = icp.Scan(ctx, data, max_dist_hint=25.0)
    with pytest.raises(icp.B200ICPError):
        d1.download(with_normals=True)                      # no normals yet
    d1.calc_normals(10, rpos)
    xyz, nrm = d1.download(with_normals=True)
    assert np.array_equal(xyz, data)
    np.testing.assert_allclose(nrm, want, rtol=0, atol=1e-12)
    m = icp.Scan(ctx, model, max_dist_hint=25.0)
    d2 = icp.Scan(ctx, data, normals=want, max_dist_hint=25.0)
    eng = icp.icp6D(ctx, algo=icp.ALGO_NAPX, max_dist_match=25.0, max_num_iterations=30, epsilon_icp=1e-5)
    r1 = eng.match(m, d1, icp.CLOSEST_PLANE_SIMPLE)
    r2 = eng.match(m, d2, icp.CLOSEST_PLANE_SIMPLE)
    assert r1["iterations"] == r2["iterations"] and np.array_equal(r1["npairs"], r2["npairs"])
    assert orclib.rel_frobenius(d1.get_pose()[0], d2.get_pose()[0]) < 1e-10
    d1.calc_normals(12, rpos)                               # replaces the normals in place
    with pytest.raises(icp.B200ICPError):
        d1.calc_normals(0, rpos)

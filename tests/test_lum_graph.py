"""LUM / graph back-end: Graph(nodes, cldist2, loopsize), FillGB3D, doGraphSlam6D.

CPU tests: the host helpers against the compiled reference's golden vectors (Matrix4ToEuler) and the oracle.
GPU tests: the whole relaxation through the C ABI against the oracle's restatement of lum6DEuler::doGraphSlam6D
(reference src/slam6d/lum6Deuler.cc:265-479), which moves points and solves the system by LU, where the product
keeps poses only and solves by Cholesky.
"""
import os

import numpy as np
import pytest

import orclib

HERE = os.path.dirname(os.path.abspath(__file__))


def test_matrix4_to_euler_matches_reference_golden(icp):
    g = np.load(os.path.join(HERE, "golden", "lum_vectors.npz"))
    for M, th, pos in zip(g["m4"], g["theta"], g["pos"]):
        p, t = icp.matrix4_to_euler(M)
        assert np.array_equal(p, pos)
        np.testing.assert_allclose(t, th, rtol=0, atol=1e-15)
        ot, op = np.zeros(3), np.zeros(3)
        orclib.port().orc_matrix4_to_euler(orclib.P(np.ascontiguousarray(M)), orclib.P(ot), orclib.P(op))
        np.testing.assert_allclose(ot, th, rtol=0, atol=1e-15)


def test_matrix4_to_euler_inverts_euler_to_matrix4(icp):
    rng = np.random.default_rng(5)
    for _ in range(50):
        pos = rng.uniform(-100, 100, 3)
        th = np.array([rng.uniform(-3, 3), rng.uniform(-1.5, 1.5), rng.uniform(-3, 3)])
        M = icp.euler_to_matrix4(pos, th)
        p, t = icp.matrix4_to_euler(M)
        assert np.array_equal(p, pos)
        # the angle triple is one of the two equivalent ones (branch on alignxf[0] > 0): compare as matrices
        np.testing.assert_allclose(icp.euler_to_matrix4(p, t), M, rtol=0, atol=1e-12)


def test_graph_from_poses(icp):
    # 8 poses on a ring of radius 300: neighbours ~230 apart, scan 0 and 7 close the loop
    ang = np.linspace(0, 2 * np.pi, 8, endpoint=False)
    rpos = np.stack([300 * np.cos(ang), np.zeros(8), 300 * np.sin(ang)], axis=1)
    g = icp.Graph.from_poses(rpos, cldist2=250.0 ** 2, loopsize=3)
    want = orclib.port_graph_from_poses(rpos, 250.0 ** 2, 3)
    assert np.array_equal(g.links, want)
    chain = [(i, i + 1) for i in range(7)]
    assert [tuple(l) for l in g.links[:7]] == chain
    assert (0, 7) in [tuple(l) for l in g.links[7:]]          # the loop closure
    assert all(k - j > 3 for j, k in g.links[7:])
    # strict '<' on the squared distance and strict '>' on the index gap (graph.cc:120-122)
    d2 = float(np.sum((rpos[0] - rpos[7]) ** 2))
    assert (0, 7) not in [tuple(l) for l in icp.Graph.from_poses(rpos, d2, 3).links]
    assert (0, 7) in [tuple(l) for l in icp.Graph.from_poses(rpos, np.nextafter(d2, np.inf), 3).links]
    assert icp.Graph.from_poses(rpos[:1], 1.0, 0).get_nr_links() == 0


def test_graph_chain(icp):
    assert [tuple(l) for l in icp.Graph.chain(4).links] == [(0, 1), (1, 2), (2, 3)]
    assert [tuple(l) for l in icp.Graph.chain(4, loop=True).links] == [(0, 1), (1, 2), (2, 3), (3, 0)]
    assert icp.Graph.chain(1).get_nr_links() == 0 and icp.Graph.chain(0).get_nr_links() == 0


def _ring_scans(icp, n_scans, n_pts, noise_pos=2.0, noise_deg=0.3, seed=0):
    """n_scans resamplings of one scene, each displaced by a small pose error the relaxation has to remove.
    Returns (list of xyz in the global frame, transMat per scan)."""
    rng = np.random.default_rng(seed)
    scans, T = [], []
    for i in range(n_scans):
        pts = icp.synth_scene(7, 100 + i, n_pts, 0.5)
        if i == 0:
            P = np.eye(4).T.reshape(16).copy()
        else:
            P = icp.euler_to_matrix4(rng.normal(0, noise_pos, 3), np.deg2rad(rng.normal(0, noise_deg, 3)))
        scans.append(icp.transform_points(P, pts))
        T.append(P)
    return scans, np.array(T)


@pytest.mark.gpu
def test_fill_gb_matches_oracle(icp, ctx):
    scans, T = _ring_scans(icp, 4, 15000)
    links = np.array([[0, 1], [1, 2], [2, 3], [0, 3]], dtype=np.int32)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, T):
        d.set_pose(t, None)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=25.0)
    G, B, npairs = lum.fill_gb(icp.Graph(links, 4), dev)
    want = orclib.port_lum_graph_slam(scans, links, 625.0, 1, -1.0, T)
    assert np.all(npairs > 1000)
    scale = np.abs(want["G"]).max()
    assert np.abs(G - want["G"]).max() <= 1e-9 * scale
    assert np.abs(B - want["B"]).max() <= 1e-9 * np.abs(want["B"]).max()
    assert np.allclose(G, G.T, rtol=0, atol=1e-12 * scale)
    # a link-sharded fill (two halves added into the same buffers) gives the same system
    G2, B2, _ = lum.fill_gb(icp.Graph(links, 4), dev, link_subset=[0, 2])
    lum.fill_gb(icp.Graph(links, 4), dev, G=G2, B=B2, link_subset=[1, 3])
    assert np.abs(G2 - G).max() <= 1e-12 * scale and np.abs(B2 - B).max() <= 1e-12 * np.abs(B).max()


@pytest.mark.gpu
@pytest.mark.parametrize("n_scans,nr_it", [(5, 1), (5, 4), (3, 3)])
def test_graph_slam_matches_oracle(icp, ctx, n_scans, nr_it):
    scans, T = _ring_scans(icp, n_scans, 15000, seed=n_scans)
    links = np.array([[i, i + 1] for i in range(n_scans - 1)] + [[0, n_scans - 1]], dtype=np.int32)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, T):
        d.set_pose(t, None)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=25.0, epsilon_lum=1e-3)
    frames = icp.Frames(n_scans)
    ret, it = lum.doGraphSlam6D(icp.Graph(links, n_scans), dev, nr_it, frames=frames)
    # transformToEuler(.., LUM, 1) for every scan but the last, (.., LUM, 2) for the last one, which also pushes a
    # frame on scan 0 (lum6Deuler.cc:447-451, scan.cc:986-999)
    assert [len(frames.get(k)) for k in range(n_scans)] == [it] * n_scans
    assert all(t == icp.FRAME_LUM for k in range(n_scans) for _, t in frames.get(k))
    assert np.array_equal(frames.get(n_scans - 1)[-1][0], dev[n_scans - 1].get_pose()[0])
    want = orclib.port_lum_graph_slam(scans, links, 625.0, nr_it, 1e-3, T)
    assert it == want["iterations"]
    assert abs(ret - want["ret"]) <= 1e-7 * max(1.0, abs(want["ret"]))
    for i, d in enumerate(dev):
        Tm, dal = d.get_pose()
        assert orclib.rel_frobenius(Tm, want["transmats"][i]) < 1e-8, i
        assert orclib.rel_frobenius(dal, want["dalignxfs"][i]) < 1e-8, i
    # scan 0 is the fixed reference
    assert np.array_equal(dev[0].get_pose()[0], T[0])
    # the relaxation moves every pose estimate towards the truth (identity: all scans sample one scene)
    if nr_it >= 3:
        for i in range(1, n_scans):
            before = np.linalg.norm(T[i][12:15])
            after = np.linalg.norm(dev[i].get_pose()[0][12:15])
            assert after < 0.75 * before, (i, before, after)


@pytest.mark.gpu
def test_graph_slam_errors(icp, ctx):
    scans, T = _ring_scans(icp, 3, 4000)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    lum = icp.lum6DEuler(ctx)
    # scan 2 is not linked to anything: G is singular -> ESTATE, poses untouched
    with pytest.raises(icp.B200ICPError) as e:
        lum.doGraphSlam6D(icp.Graph(np.array([[0, 1]], dtype=np.int32), 3), dev, 1)
    assert e.value.code == -6
    with pytest.raises(icp.B200ICPError):
        lum.doGraphSlam6D(icp.Graph(np.array([[0, 5]], dtype=np.int32), 3), dev, 1)


def _golden_link_inputs():
    rng = np.random.default_rng(12)
    base = rng.uniform(-300, 300, (8000, 3)); base[:, 1] = np.abs(base[:, 1]) * 0.3
    model = np.ascontiguousarray(base + rng.normal(0, 0.3, base.shape))
    data = np.ascontiguousarray(base[:6000] + rng.normal(0, 0.3, (6000, 3)) + [1.0, -0.5, 0.7])
    return model, data


def test_oracle_lum_link_matches_reference_golden():
    """C / CD of one link against the stored output of the compiled reference's getPtPairs + newmat (works where
    oracle/_ref is absent; the live variant is tests/test_oracle_pinning.py)"""
    g = np.load(os.path.join(HERE, "golden", "lum_vectors.npz"))
    model, data = _golden_link_inputs()
    C, CD, m = orclib.port_lum_link(model, data, 100.0, g["link_S"])
    assert m == int(g["link_pairs"][0]) and m > 1000
    np.testing.assert_allclose(C.reshape(-1), g["link_C"], rtol=1e-9)
    np.testing.assert_allclose(CD, g["link_CD"], rtol=1e-9, atol=1e-9 * np.abs(g["link_CD"]).max())


@pytest.mark.gpu
def test_gpu_lum_link_matches_reference_golden(icp, ctx):
    g = np.load(os.path.join(HERE, "golden", "lum_vectors.npz"))
    model, data = _golden_link_inputs()
    first, second = icp.Scan(ctx, model, max_dist_hint=10.0), icp.Scan(ctx, data, max_dist_hint=10.0)
    first.set_pose(None, g["link_S"])                      # Source->dalignxf
    for _ in range(2):                                     # second evaluation runs seeded from the first: same result
        C, CD, m = icp.lum_link(ctx, first, second, 100.0)
        assert m == int(g["link_pairs"][0])
        np.testing.assert_allclose(C.reshape(-1), g["link_C"], rtol=1e-9)
        np.testing.assert_allclose(CD, g["link_CD"], rtol=1e-9, atol=1e-9 * np.abs(g["link_CD"]).max())

"""Shared inputs of tests/test_full_reference.py and tests/golden/make_full_golden.py (the reference's own Scan /
icp6D / lum6DEuler / BOctTree classes, oracle/full_harness.cc)."""
import numpy as np

LUM_SCANS, LUM_PTS = 5, 8000
LUM_LINKS = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [0, 4], [1, 3]], dtype=np.int32)
LUM_PARAMS = dict(nr_it=4, max_dist_lum=25.0, eps_lum=1e-9)
OCT_VOXELS = (10.0, 3.0)


def lum_sequence(icp):
    """Scan-local clouds (independent samplings of one scene seen from drifting poses) + their poses."""
    rng = np.random.default_rng(2024)
    locals_, rpos, rtheta = [], [], []
    for i in range(LUM_SCANS):
        p = np.array([40.0 * i, 0.0, 15.0 * i]) + rng.normal(0, 3.0, 3) * (i > 0)
        t = np.deg2rad(np.array([0.0, 4.0 * i, 0.0]) + rng.normal(0, 0.4, 3) * (i > 0))
        true_p = np.array([40.0 * i, 0.0, 15.0 * i])
        true_t = np.deg2rad(np.array([0.0, 4.0 * i, 0.0]))
        world = icp.synth_scene(7, 500 + i, LUM_PTS, 0.5)
        Minv, _ = icp.m4inv(icp.euler_to_matrix4(true_p, true_t))
        locals_.append(icp.transform_points(Minv, world))      # what a scanner at the TRUE pose records
        rpos.append(p); rtheta.append(t)                        # ... loaded with a drifted pose estimate
    return locals_, np.array(rpos), np.array(rtheta)


def cov_pair(icp):
    model = icp.synth_scene(7, 42, 20000, 0.5)
    data = icp.synth_scene(7, 43, 15000, 0.5)
    return model, data, np.array([3.0, -2.0, 1.5]), np.deg2rad(np.array([0.2, -0.3, 0.25]))


# points ON splitting planes of the octree: bbox [0,4]^3 -> root centre (2,2,2), half-size 3
OCT_PLANE_KAT = np.array([[0, 0, 0], [4, 4, 4], [2, 2, 2], [2, 1, 3], [1, 2, 3]], dtype=np.float64)

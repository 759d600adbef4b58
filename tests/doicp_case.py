"""Shared input of the doICP parity tests and of tests/golden/make_doicp_golden.py: a short scan sequence with
accumulating odometry drift (every scan is an independent sampling of one scene, loaded at a drifting pose)."""
import numpy as np

N_SCANS, N_PTS = 4, 20000
PARAMS = dict(algo=1, max_dist=25.0, max_iter=30, eps=1e-5)
VARIANTS = [(True, False, 0), (False, False, 0), (True, True, 0), (True, True, 2)]   # (eP, meta, max_num_metascans)


def make_sequence(icp):
    rng = np.random.default_rng(77)
    scans, org = [], []
    E = np.eye(4)
    for i in range(N_SCANS):
        if i > 0:
            d = icp.euler_to_matrix4(rng.normal(0, 2.5, 3), np.deg2rad(rng.normal(0, 0.25, 3))).reshape(4, 4).T
            E = E @ d
        M = E.T.reshape(16).copy()
        scans.append(icp.transform_points(M, icp.synth_scene(7, 300 + i, N_PTS, 0.5)))
        org.append(M)
    return scans, np.array(org)

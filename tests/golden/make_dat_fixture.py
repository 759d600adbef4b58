#!/usr/bin/env python
"""Builds tests/golden/dat_reduced.npz from the reference's bundled scans (BASELINE configs[0]:
`bin/slam6D -r 10 -i 20 dat`): reads dat/scan00{0,1,2}.{3d,pose} (uos format: `x y z` per line; pose = position +
Euler angles in DEGREES, src/scanio/helper.cc:228-232), applies the octree voxel-CENTRE reduction the reference
applies for `-r 10` (restated from include/slam6d/Boctree.h:224-270, :612-656, :928-949, :1164-1195, :1353-1355:
root cube = bbox centre, half-size = max half-extent + 1.0; child index bit k set iff p[k] > centre[k]; a child
is a leaf when ITS half-size <= voxel; output = leaf-cube centres, depth first, children 0..7), moves the reduced
points by the scan's pose (BasicScan::calcReducedOnDemandPrivate, basicScan.cc:730-737) and stores them together
with the results of the compiled reference (oracle/_ref) on them.  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib  # noqa: E402
from orclib import P  # noqa: E402

DAT = "/root/reference/dat"


octree_centres = orclib.octree_centres


def main():
    L = orclib.ref()
    assert L is not None
    out = {}
    scans = []
    for k in range(3):
        pts = np.loadtxt(os.path.join(DAT, "scan%03d.3d" % k))[:, :3]
        pose = np.loadtxt(os.path.join(DAT, "scan%03d.pose" % k)).reshape(2, 3)
        M = np.empty(16)
        L.ref_euler_to_matrix4(P(np.ascontiguousarray(pose[0])), P(np.deg2rad(pose[1])), P(M))
        red = octree_centres(pts, 10.0)
        Mm = M.reshape(4, 4).T
        red = np.ascontiguousarray(red @ Mm[:3, :3].T + Mm[:3, 3])      # transformReduced(transMatOrg)
        out["scan%d_raw_count" % k] = np.array([len(pts)])
        if k == 1:
            out["scan1_raw_first30000"] = np.ascontiguousarray(pts[:30000])   # input of the reduction parity test
        out["scan%d_transMatOrg" % k] = M
        out["scan%d_xyz_reduced" % k] = red
        scans.append(red)
        print("scan", k, len(pts), "->", len(red), "reduced points")
    # sequential matching scan0 <- scan1 <- scan2 (icp6D::doICP order, no pose extrapolation), QUAT, -i 20, d = 25
    dal = orclib.identity()
    for k in (1, 2):
        r = orclib.ref_match(scans[k - 1], scans[k], algo=1, max_dist=25.0, max_iter=20, eps=1e-5, model_dalignxf=dal)
        out["match%d_transmat" % k] = r["transmat"]
        out["match%d_rms" % k] = r["rms"]
        out["match%d_npairs" % k] = r["npairs"]
        out["match%d_iterations" % k] = np.array([r["iterations"]])
        dal = r["dalignxf"]
        print("match", k, "iterations", r["iterations"], "pairs", r["npairs"][-1], "rms", r["rms"][-1])
    path = os.path.join(HERE, "dat_reduced.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

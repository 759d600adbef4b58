#!/usr/bin/env python
"""Golden vectors for the LUM / graph back-end helpers, generated from the COMPILED REFERENCE
(oracle/_ref/libref3dtk.so, built from /root/reference by oracle/Makefile).  Run in the build container:
    python tests/golden/make_lum_golden.py      -> tests/golden/lum_vectors.npz
Contents: Matrix4ToEuler (globals.icc:540-578) on poses covering both asin branches and the gimbal-lock branch.
doGraphSlam6D, covarianceEuler / covarianceQuat and Graph() are pinned separately, by the reference's own classes
(tests/golden/make_full_golden.py, oracle/_ref/libref3dtk_full.so)."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orclib

L = orclib.ref()
assert L is not None, "build oracle/_ref first (make -C oracle ref)"
L.ref_matrix4_to_euler.restype = None
L.ref_matrix4_to_euler.argtypes = [orclib.vp, orclib.vp, orclib.vp]
rng = np.random.default_rng(20261017)
poses, mats, thetas, positions = [], [], [], []
cases = [np.array([0.0, 0.0, 0.0]), np.array([0.3, -0.2, 0.1]), np.array([0.1, 2.5, -0.4]),      # cos(ty) < 0 branch
         np.array([-1.2, -2.9, 3.0]), np.array([0.4, np.pi / 2, 0.2]), np.array([0.4, -np.pi / 2 + 1e-4, 1.0])]
cases += [rng.uniform(-np.pi, np.pi, 3) for _ in range(58)]
for th in cases:
    pos = rng.uniform(-500, 500, 3)
    M = np.zeros(16)
    L.ref_euler_to_matrix4(orclib.P(pos), orclib.P(np.asarray(th, dtype=np.float64)), orclib.P(M))
    out_th, out_pos = np.zeros(3), np.zeros(3)
    L.ref_matrix4_to_euler(orclib.P(M), orclib.P(out_th), orclib.P(out_pos))
    mats.append(M); thetas.append(out_th); positions.append(out_pos)
# LUM link (covarianceEuler): pairs by the compiled reference's KDtree / getPtPairs, sums restated in the harness with
# the reference's newmat inverse (ref_lum_link; lum6Deuler.cc itself does not compile here).  Inputs are regenerated
# from the seed by the test, only the outputs are stored.
lrng = np.random.default_rng(12)
base = lrng.uniform(-300, 300, (8000, 3)); base[:, 1] = np.abs(base[:, 1]) * 0.3
lmodel = np.ascontiguousarray(base + lrng.normal(0, 0.3, base.shape))
ldata = np.ascontiguousarray(base[:6000] + lrng.normal(0, 0.3, (6000, 3)) + [1.0, -0.5, 0.7])
S = np.empty(16)
L.ref_euler_to_matrix4(orclib.P(np.array([0.4, 0.2, -0.3])), orclib.P(np.deg2rad([0.1, -0.2, 0.15])), orclib.P(S))
rt = L.ref_tree_create(orclib.P(lmodel), len(lmodel), 0, 20)
linkC, linkCD = np.zeros(36), np.zeros(6)
linkm = L.ref_lum_link(rt, orclib.P(S), orclib.P(ldata), len(ldata), 100.0, orclib.P(linkC), orclib.P(linkCD))
L.ref_tree_free(rt)
np.savez_compressed(os.path.join(HERE, "lum_vectors.npz"), m4=np.array(mats), theta=np.array(thetas),
                    pos=np.array(positions), link_C=linkC, link_CD=linkCD, link_pairs=np.array([linkm]), link_S=S)
print("wrote", len(mats), "Matrix4ToEuler vectors")

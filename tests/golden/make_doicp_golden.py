#!/usr/bin/env python
"""Golden results of icp6D::doICP on tests/doicp_case.py's sequence, every match run by the COMPILED REFERENCE
(oracle/_ref/libref3dtk.so: kd.cc, searchTree.cc, icp6Dquat.cc unmodified; the doICP / Scan::transform glue is
orclib.do_icp).  Run in the build container:  python tests/golden/make_doicp_golden.py -> doicp_vectors.npz"""
import importlib, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.dirname(HERE)]
import orclib, doicp_case

icp = importlib.import_module("3dtk_b200")     # host helpers only (scene generator, EulerToMatrix4)
assert orclib.ref() is not None, "build oracle/_ref first (make -C oracle ref)"
scans, org = doicp_case.make_sequence(icp)
out = {}
for eP, meta, mx in doicp_case.VARIANTS:
    r = orclib.do_icp(orclib.ref_match, scans, org, extrapolate_pose=eP, meta=meta, max_num_metascans=mx,
                      **doicp_case.PARAMS)
    key = "eP%d_meta%d_max%d" % (eP, meta, mx)
    out[key + "_transmats"] = r["transmats"]
    out[key + "_iterations"] = np.array(r["iterations"])
    print(key, r["iterations"])
np.savez_compressed(os.path.join(HERE, "doicp_vectors.npz"), **out)

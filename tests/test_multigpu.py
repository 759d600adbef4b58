"""Query-sharded single match across GPUs: every rank holds the model and a slice of the data scan;
the moment all-reduce is fused into the iteration kernel over NVLink peer memory.  Needs >= 2 GPUs in the box
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`); skipped on a single-GPU box."""
import threading

import numpy as np
import pytest

import orclib
from conftest import make_pair

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_query_sharded_match_equals_single_gpu_and_oracle(icp, world):
    if _ngpus() < world:
        pytest.skip("needs %d GPUs" % world)
    model, data, _ = make_pair(icp, 60000, 50001)          # ragged split on purpose
    want = orclib.port_match(model, data, algo=1, max_dist=25.0, max_iter=50, eps=1e-5)
    ctxs = [icp.Context(r) for r in range(world)]
    for r, c in enumerate(ctxs):
        c.comm_create(r, world)
    icp.Context.comm_connect_local(ctxs)
    step = -(-len(data) // world)                           # ceil, like getPtPairsParallel (scan.cc:1335-1342)
    out = [None] * world

    def rank_main(r):
        c = ctxs[r]
        m = icp.Scan(c, model, max_dist_hint=25.0)
        d = icp.Scan(c, data[r * step:min((r + 1) * step, len(data))], max_dist_hint=25.0)
        res = icp.icp6D(c, algo=icp.ALGO_QUAT, max_dist_match=25.0, max_num_iterations=50, epsilon_icp=1e-5,
                        sharded=True).match(m, d)
        out[r] = (res, d.get_pose()[0].copy())

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert all(o is not None for o in out), "a rank did not finish"
    T0 = out[0][1]
    for r in range(world):
        res, T = out[r]
        assert np.array_equal(T, T0)                        # bit-identical loop state on every rank
        assert res["iterations"] == want["iterations"]
        np.testing.assert_allclose(res["rms"], want["rms"], rtol=1e-9)
    assert sum(int(out[r][0]["npairs"][-1]) for r in range(1)) > 0
    assert int(out[0][0]["npairs"][-1]) == int(want["npairs"][-1])   # npairs is the GLOBAL count (summed moments)
    assert orclib.rel_frobenius(T0, want["transmat"]) < 1e-8
    for c in ctxs:
        c.close()

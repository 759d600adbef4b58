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


# ---- link-sharded LUM: one process per GPU, NCCL all-reduce of the global [G|B] only -------------
def _lum_rank(rank, world, port, q):
    import importlib
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    icp = importlib.import_module("3dtk_b200")
    par = importlib.import_module("3dtk_b200.parallel")
    import test_lum_graph
    scans, T = test_lum_graph._ring_scans(icp, 5, 15000, seed=5)
    links = np.array([[i, i + 1] for i in range(4)] + [[0, 4], [1, 3]], dtype=np.int32)
    ctx = icp.Context(rank)
    dev = [icp.Scan(ctx, s, max_dist_hint=25.0) for s in scans]
    for d, t in zip(dev, T):
        d.set_pose(t, None)
    lum = icp.lum6DEuler(ctx, max_dist_match_lum=25.0, epsilon_lum=1e-3)
    ret, it = par.graph_slam_sharded(lum, icp.Graph(links, 5), dev, 4, rank, world, device=torch.device("cuda", rank))
    q.put((rank, ret, it, np.array([d.get_pose()[0] for d in dev])))
    dist.barrier()
    dist.destroy_process_group()


def test_link_sharded_graph_slam_nccl(icp):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import socket
    import torch.multiprocessing as mp
    import test_lum_graph
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_lum_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    scans, T = test_lum_graph._ring_scans(icp, 5, 15000, seed=5)
    links = np.array([[i, i + 1] for i in range(4)] + [[0, 4], [1, 3]], dtype=np.int32)
    want = orclib.port_lum_graph_slam(scans, links, 625.0, 4, 1e-3, T)
    assert np.array_equal(res[0][3], res[1][3])              # replicated poses stay bit-identical across ranks
    for rank, ret, it, poses in res:
        assert it == want["iterations"] and abs(ret - want["ret"]) < 1e-7
        for i in range(5):
            assert orclib.rel_frobenius(poses[i], want["transmats"][i]) < 1e-8

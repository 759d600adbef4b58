"""N > 1 host logic on CPU: world_size-2 gloo processes shard graph links, fill their part of the LUM system
and all-reduce it; the result must equal the serial FillGB3D.  Also the bench timing reduction."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _links_and_blocks(n_scans, seed=0):
    rng = np.random.default_rng(seed)
    links = [(i - 1, i) for i in range(1, n_scans)] + [(0, n_scans - 1), (2, n_scans - 2)]   # chain + two loops
    blocks = []
    for _ in links:
        A = rng.normal(size=(6, 6))
        blocks.append((A @ A.T + 6 * np.eye(6), rng.normal(size=6)))
    return links, blocks


def _worker(rank, world, port, n_scans, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module("3dtk_b200.parallel")
    links, blocks = _links_and_blocks(n_scans)
    mine = par.shard_units(len(links), rank, world)
    G, B = par.fill_gb([links[i] for i in mine], [blocks[i] for i in mine], n_scans)
    G, B = par.allreduce_gb(G, B)
    t, u = par.reduce_timing(1.0 + rank, 10 * (rank + 1))
    q.put((rank, G, B, t, u, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_link_sharded_lum_system_matches_serial_fill():
    par = importlib.import_module("3dtk_b200.parallel")
    n_scans, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scans, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    links, blocks = _links_and_blocks(n_scans)
    Gs, Bs = par.fill_gb(links, blocks, n_scans)
    seen = []
    for rank, G, B, t, u, mine in res:
        np.testing.assert_allclose(G, Gs, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(B, Bs, rtol=1e-13, atol=1e-12)
        assert t == 2.0 and u == 30.0             # max over ranks, sum of units
        seen += mine
    assert sorted(seen) == list(range(len(links)))  # every link on exactly one rank
    # the assembled system is what the reference solves: symmetric positive definite here
    assert np.allclose(Gs, Gs.T) and np.linalg.eigvalsh(Gs).min() > 0


def test_sharding_and_pose_chain():
    par = importlib.import_module("3dtk_b200.parallel")
    assert par.shard_units(10, 1, 4) == [1, 5, 9]
    assert par.shard_units(3, 3, 4) == []
    with pytest.raises(ValueError):
        par.shard_units(3, 4, 4)
    assert par.sequential_pairs(4) == [(0, 1), (1, 2), (2, 3)]
    icp = importlib.import_module("3dtk_b200")
    rel = [icp.euler_to_matrix4(np.array([1.0 * k, 0.5, -0.2]), np.array([0.01 * k, 0.0, 0.02])) for k in (1, 2, 3)]
    poses = par.chain_poses(rel)
    want = icp.mmult(rel[2], icp.mmult(rel[1], rel[0]))
    np.testing.assert_allclose(poses[3], want, atol=1e-13)
    assert np.array_equal(poses[0], np.eye(4).reshape(16))

"""Multi-GPU host logic for the path: how units are sharded and the one exchange step.

The north-star shards *scan pairs* (sequential ICP without metascan) and *graph links* (LUM FillGB3D) across
ranks -- one process per GPU, no collective on the data path.  The only exchange is the global LUM system:
every rank adds the blocks of its links into a dense (6n x 6n) G and a 6n-vector B, then ONE all-reduce of
the packed buffer makes them global (reference: the `#pragma omp critical` accumulation of
lum6DEuler::FillGB3D, src/slam6d/lum6Deuler.cc:285-298).

Backend-agnostic: `torch.distributed` with "nccl" on the GPU box, "gloo" in the CPU tests.
"""
import numpy as np


def shard_units(n_units, rank, world_size):
    """Round-robin assignment of independent units (scan pairs, graph links) to ranks.
    Round-robin rather than contiguous: consecutive scan pairs have similar cost, so every rank gets a
    similar mix (the reference uses `schedule(dynamic)` over links, lum6Deuler.cc:271)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    return list(range(rank, n_units, world_size))


def sequential_pairs(n_scans):
    """(previous, current) pairs of icp6D::doICP without metascan (icp6D.cc:374-437): independent units,
    because the pose extrapolation applies the same delta to both scans of a pair."""
    return [(i - 1, i) for i in range(1, n_scans)]


def chain_poses(relative):
    """Prefix product of the per-pair 4x4 (column-major 16-vectors) results: pose_i = rel_i * pose_{i-1}.
    The host-side step that follows pair-sharded matching."""
    out = [np.eye(4)]
    for r in relative:
        out.append(np.asarray(r, dtype=np.float64).reshape(4, 4).T @ out[-1])
    return [m.T.reshape(16).copy() for m in out]


def fill_gb(links, blocks, n_scans):
    """FillGB3D for a subset of links.  links: [(first, second)] scan numbers with scan 0 fixed;
    blocks: [(C[6,6], CD[6])] from b200icp_lum_link.  Returns dense G ((n-1)*6 square) and B."""
    dim = 6 * (n_scans - 1)
    G = np.zeros((dim, dim))
    B = np.zeros(dim)
    for (first, second), (C, CD) in zip(links, blocks):
        a, b = first - 1, second - 1                      # lum6Deuler.cc:273-274
        C = np.asarray(C, dtype=np.float64).reshape(6, 6)
        CD = np.asarray(CD, dtype=np.float64).reshape(6)
        if a >= 0:
            B[6 * a:6 * a + 6] += CD
            G[6 * a:6 * a + 6, 6 * a:6 * a + 6] += C
        if b >= 0:
            B[6 * b:6 * b + 6] -= CD
            G[6 * b:6 * b + 6, 6 * b:6 * b + 6] += C
        if a >= 0 and b >= 0:
            G[6 * a:6 * a + 6, 6 * b:6 * b + 6] -= C
            G[6 * b:6 * b + 6, 6 * a:6 * a + 6] -= C
    return G, B


def allreduce_gb(G, B, group=None, device=None):
    """One all-reduce (SUM) of the packed [G | B] fp64 buffer -> the global LUM system on every rank."""
    import torch
    import torch.distributed as dist
    dim = B.shape[0]
    buf = torch.empty(dim * dim + dim, dtype=torch.float64, device=device)
    buf[:dim * dim] = torch.from_numpy(np.ascontiguousarray(G).reshape(-1))
    buf[dim * dim:] = torch.from_numpy(np.ascontiguousarray(B))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    out = buf.cpu().numpy()
    return out[:dim * dim].reshape(dim, dim).copy(), out[dim * dim:].copy()


def reduce_timing(elapsed_s, units, group=None, device=None):
    """bench.py contract: time = MAX over ranks, work = SUM over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(elapsed_s)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(u, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(u.item())


def graph_slam_sharded(lum, graph, scans, nr_it, rank, world_size, group=None, device=None, frames=None):
    """lum6DEuler::doGraphSlam6D with the links of FillGB3D sharded over ranks (the north-star's
    "NCCL allreduce only for the global lum6D covariance sum").  The loop itself is the C ABI's
    `b200icp_lum_graph_slam_sharded` (what a C++ host calls with an MPI / NCCL callback); this wrapper only supplies
    the all-reduce: ONE `torch.distributed.all_reduce` of the packed [G|B] fp64 buffer per LUM iteration ("nccl" on
    the GPU box -- staged through `device` --, "gloo" in the CPU tests).  Returns (ret, iterations)."""
    import torch
    import torch.distributed as dist
    if len(scans) <= 0:
        raise ValueError("Zero scans in graph")            # lum6Deuler.cc:316-318

    def allreduce(buf):
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        t = torch.from_numpy(buf)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        buf[:] = t.cpu().numpy()

    return lum.doGraphSlam6D_sharded(graph, scans, nr_it, rank, world_size, allreduce, frames=frames)

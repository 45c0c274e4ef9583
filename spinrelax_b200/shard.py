"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed for
the little communication there is.

  * C(t), histogram, fits, relaxation: shard by bond vector / residue (independent units) -- no data-path
    collective; the per-vector result rows are gathered to rank 0.
  * dq moments: shard by lag (every rank holds the 16 MB quaternion array) and gather, or shard by frame range /
    replica trajectory and all-reduce the raw FP64 moment sums.
  * histogram sharded by frames (independent trajectories): one all-reduce (sum) of the integer counts.
The functions work with any initialised process group (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def split_range(n, world, rank):
    """Contiguous balanced partition of range(n): the first n % world ranks get one extra item."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def split_sizes(n, world):
    return [split_range(n, world, r)[1] - split_range(n, world, r)[0] for r in range(world)]


def shard_vectors(vecs, world, rank):
    """vecs (..., nR, 3) -> this rank's contiguous block of bond vectors (a view)."""
    a, b = split_range(vecs.shape[-2], world, rank)
    return vecs[..., a:b, :]


def shard_lags(lags, world, rank):
    """Round-robin over the lag list so that every rank gets a similar mix of short and long lags."""
    return np.asarray(lags)[rank::world]


def gather_columns(local, n_total, dst=0):
    """Gather per-vector result columns: local (rows, n_local) on every rank -> (rows, n_total) on `dst`
    (None elsewhere).  Column blocks follow split_range."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = split_sizes(n_total, world)
    if local.shape[1] != sizes[rank]:
        raise ValueError("gather_columns: rank %d holds %d columns, the partition of %d over %d ranks gives it %d"
                         % (rank, local.shape[1], n_total, world, sizes[rank]))
    rows = local.shape[0]
    width = max(sizes)
    pad = torch.zeros((rows, width), dtype=local.dtype, device=local.device)
    pad[:, : local.shape[1]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:, : sizes[r]] for r in range(world)], dim=1)


def gather_rows(local, n_total, dst=0):
    """Gather per-vector blocks along axis 0: local (n_local, ...) on every rank -> (n_total, ...) on `dst` (None
    elsewhere).  Row blocks follow split_range (e.g. the (nR, nbx, nby) histogram counts)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = split_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError("gather_rows: rank %d holds %d rows, the partition of %d over %d ranks gives it %d"
                         % (rank, local.shape[0], n_total, world, sizes[rank]))
    pad = torch.zeros((max(sizes),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)], dim=0)


def merge_lag_results(local, lags, world, dst=0):
    """Inverse of shard_lags: local (n_local, ...) rows for lags[rank::world] -> (nLags, ...) on `dst`."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    n = len(lags)
    width = -(-n // world)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    out = torch.empty((n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        k = len(range(r, n, world))
        out[r::world] = bufs[r][:k]
    return out


def allreduce_sum_(t):
    """In-place sum over ranks (histogram counts / raw dq moment sums when sharded by frames or replicas)."""
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t

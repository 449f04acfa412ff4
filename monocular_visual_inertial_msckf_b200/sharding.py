"""Multi-GPU plumbing of the Monte-Carlo batch: filter instances are independent, so the filter
axis is cut into contiguous shards, one per rank / GPU, with NO data-path collective; the only exchange is a final gather
of per-filter results (trajectories, error statistics) to rank 0 over torch.distributed (NCCL on the GPU box, gloo in the
CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_filters: int, rank: int, world_size: int):
    """Contiguous [begin, end) of the filters rank `rank` owns; sizes differ by at most one."""
    base, rem = divmod(int(n_filters), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_filter_results(local: torch.Tensor, n_filters: int, dst: int = 0):
    """Concatenates the per-rank [n_local, ...] result tensors in filter order on rank `dst` (None elsewhere).
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(n_filters, r, world) for r in range(world)]
    n_max = max(e - b for b, e in sizes)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:e - b] for r, (b, e) in enumerate(sizes)], dim=0)


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the maximum over ranks."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

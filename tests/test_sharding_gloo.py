"""CPU, world_size 2 over gloo: the filter axis is sharded contiguously and the final gather restores filter order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from monocular_visual_inertial_msckf_b200.sharding import gather_filter_results, max_over_ranks, shard_range


def test_shard_range_partitions_the_batch():
    for n in (1, 2, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_filters, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(n_filters, rank, world)
    # stand-in for the per-filter result of this rank's shard (position of filter i after the run): a function of i only
    idx = torch.arange(b, e, dtype=torch.float64)
    local = torch.stack([idx, 2 * idx + 1, idx * idx], dim=1)
    full = gather_filter_results(local, n_filters, dst=0)
    t = max_over_ranks(10.0 + rank)
    if rank == 0:
        ref = torch.arange(n_filters, dtype=torch.float64)
        ok = full is not None and torch.equal(full, torch.stack([ref, 2 * ref + 1, ref * ref], dim=1)) and t == 10.0 + world - 1
        out.put(bool(ok))
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_filters", [8, 9])
def test_gather_restores_filter_order_world2(n_filters):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_filters, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True

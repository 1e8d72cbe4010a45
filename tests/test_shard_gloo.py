"""CPU, world_size 2 over gloo: the multi-GPU path's host logic (contiguous sharding + ordered gather of maps)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xfr_b200.shard import gather_maps, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 1023):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n and hi - lo in (n // world, n // world + 1)
                cover += list(range(lo, hi))
            assert cover == list(range(n))


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    # stand-in for the engine: the "map" of triplet i is filled with i (+ a position ramp)
    local = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1, 1) + torch.linspace(0, 0.5, 12).view(1, 3, 4)
    out = gather_maps(local, n, dst=0)
    if rank == 0:
        want = torch.arange(n, dtype=torch.float32).view(-1, 1, 1) + torch.linspace(0, 0.5, 12).view(1, 3, 4)
        q.put(bool(torch.equal(out, want)))
    else:
        assert out is None
    dist.destroy_process_group()


def test_gather_in_triplet_order_world2():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    for n in (8, 7):                      # even and ragged shards
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert q.get(timeout=10) is True

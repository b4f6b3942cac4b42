"""CPU, world_size 2/3 (gloo): the host-side logic of the multi-GPU dense extraction —
slab partition, all-gather of logit slabs (even and ragged), min/max key exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nx, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vtaco_b200 import dist as vd
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(nx ** 3, dtype=torch.float32).reshape(nx, nx, nx) * 0.5 - 7.0
        grid = torch.full((nx, nx, nx), float('nan'))
        x0, x1 = vd.slab(nx, rank, world)
        grid[x0:x1] = full[x0:x1]
        vd.all_gather_slabs(grid, nx)
        ok_grid = torch.equal(grid, full)
        # min/max keys: use plain ints that behave like the ordered-int keys
        lo, hi = int(full[x0:x1].min().item() * 2) if x1 > x0 else 2 ** 31 - 1, \
            int(full[x0:x1].max().item() * 2) if x1 > x0 else -2 ** 31
        keys = torch.tensor([lo, hi], dtype=torch.int32)
        vd.all_reduce_minmax(keys)
        ok_keys = keys.tolist() == [int(full.min().item() * 2), int(full.max().item() * 2)]
        q.put((rank, ok_grid, ok_keys, (x0, x1)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,nx', [(2, 8), (2, 7), (3, 8), (3, 2)])
def test_slab_allgather_and_minmax(world, nx):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    slabs = sorted(r[3] for r in res)
    assert all(r[1] and r[2] for r in res), res
    # slabs tile [0, nx) without overlap
    cover = []
    for a, b in slabs:
        cover += list(range(a, b))
    assert cover == list(range(nx))


def test_slab_partition_properties():
    from vtaco_b200.dist import slab
    for nx in (1, 7, 8, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            rows = [slab(nx, r, world) for r in range(world)]
            assert rows[0][0] == 0 and max(b for _, b in rows) == nx
            assert sum(b - a for a, b in rows) == nx
            per = -(-nx // world)
            assert all(b - a <= per for a, b in rows)
    assert slab(256, 3, 8) == (96, 128)


def test_single_process_is_identity():
    from vtaco_b200 import dist as vd
    g = torch.randn(4, 4, 4)
    assert vd.rank_world(None) == (0, 1)
    assert vd.all_gather_slabs(g.clone(), 4) .equal(g)
    k = torch.tensor([3, 9], dtype=torch.int32)
    assert vd.all_reduce_minmax(k.clone()).equal(k)

"""GPU, >= 2 devices (NCCL): slab-sharded dense extraction reproduces the 1-GPU logit grid
bit-for-bit and the same mesh."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev, nx):
    from vtaco_b200.conv_onet.models import decoder_dict, ConvolutionalOccupancyNetwork
    from vtaco_b200.conv_onet.generation import Generator3D
    torch.manual_seed(0)
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32)
    with torch.no_grad():
        for b in dec.blocks:
            b.fc_1.weight.normal_(0, 0.1)
    net = ConvolutionalOccupancyNetwork(dec, None, device=dev)
    gen = Generator3D(net, device=dev, resolution0=nx // 4, with_img=False, padding=0.1, input_type='pointcloud')
    g = torch.Generator().manual_seed(1)
    c = {'grid': torch.randn(1, 32, 32, 32, 32, generator=g).to(dev)}
    return gen, c


def _worker(rank, world, port, nx, q, exchange='nccl'):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        gen, c = _build(dev, nx)
        if exchange == 'root':
            _root_worker(gen, c, nx, rank, q)
            return
        if exchange == 'mesh':
            _mesh_worker(gen, c, nx, rank, q)
            return
        grid, keys = gen.eval_lattice(c, group=dist.group.WORLD, exchange=exchange)
        v, f = gen.extract_mesh(grid, keys)
        v, f = v.clone(), f.clone()
        single, skeys = gen.eval_lattice(c, group=False)   # this rank alone, whole lattice
        single, skeys = single.clone(), skeys.clone()   # the generator re-uses these buffers
        from vtaco_b200.mcubes import keys_to_level
        lvl = keys_to_level(keys)
        grid2, _ = gen.eval_lattice(c, group=dist.group.WORLD, exchange=exchange)
        ok = torch.equal(grid2, single) and lvl == keys_to_level(skeys)
        v1, f1 = gen.extract_mesh(single, skeys)
        ok_mesh = torch.equal(v, v1) and torch.equal(f, f1)
        q.put((rank, bool(ok), bool(ok_mesh), int(f.shape[0])))
    finally:
        dist.destroy_process_group()


def _mesh_worker(gen, c, nx, rank, q):
    """sharded marching cubes + gather of mesh pieces: eager (with buffer growth), then CUDA-graph
    replays; the assembled mesh must equal the single-GPU mesh bit-for-bit (ids and order too)."""
    single, skeys = gen.eval_lattice(c, group=False)
    v1, f1 = [t.clone() for t in gen.extract_mesh(single.clone(), skeys.clone())]
    ok, ok_mesh = True, True
    for gather in ('root', 'all'):
        gen.mesh_gather = gather
        res = gen.generate_mesh(c=c, group=dist.group.WORLD, exchange='mesh', to_host=False)
        if rank == 0 or gather == 'all':
            ok = ok and res is not None and torch.equal(res[0], v1) and torch.equal(res[1], f1)
        else:
            ok = ok and res is None
        graph, out = gen.capture_step(c, group=dist.group.WORLD, exchange='mesh')
        for _ in range(4):
            graph.replay()
        torch.cuda.synchronize()
        V, F = [int(x) for x in out[2].cpu()]
        ok_mesh = ok_mesh and (V, F) == (v1.shape[0], f1.shape[0])
        if rank == 0 or gather == 'all':
            ok_mesh = ok_mesh and torch.equal(out[0][:V], v1) and torch.equal(out[1][:F], f1)
        ok_mesh = ok_mesh and not gen._mesh_ex.timed_out()
        dist.barrier()
    q.put((rank, bool(ok), bool(ok_mesh), int(f1.shape[0])))


def _root_worker(gen, c, nx, rank, q):
    """gather-to-root, double-buffered: several steps eager, then through the two alternating graphs."""
    gen.root_rows = max(2, nx // dist.get_world_size() - 6)
    single, skeys = gen.eval_lattice(c, group=False)
    single, skeys = single.clone(), skeys.clone()
    v1, f1 = [t.clone() for t in gen.extract_mesh(single, skeys)]
    ok = True
    for _ in range(3):
        grid, keys = gen.eval_lattice(c, group=dist.group.WORLD, exchange='root')
        if rank == 0:
            ok = ok and torch.equal(grid, single)
            v, f = gen.extract_mesh(grid, keys)
            ok = ok and torch.equal(v, v1) and torch.equal(f, f1)
        else:
            ok = ok and grid is None
    if gen._root_ex.parity:          # leave the eager steps on an even count before capturing
        gen.eval_lattice(c, group=dist.group.WORLD, exchange='root')
    stepper, out = gen.capture_step(c, group=dist.group.WORLD, exchange='root')
    for _ in range(5):
        stepper.replay()
    torch.cuda.synchronize()
    ok_mesh = True
    if rank == 0:
        V, F = [int(x) for x in out[2][:2].cpu()]
        from vtaco_b200.mcubes import keys_to_level
        ok_mesh = F == f1.shape[0] and torch.equal(out[1][:F], f1)
        nxh = np.float32(nx / 2)
        ok_mesh = ok_mesh and V == v1.shape[0]
    dist.barrier()
    q.put((rank, bool(ok), bool(ok_mesh), int(f1.shape[0])))


@pytest.mark.parametrize('exchange', ['mesh', 'fused', 'nccl', 'root'])
@pytest.mark.parametrize('nx', [64, 40])
def test_sharded_extraction_matches_single_gpu(nx, exchange):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip('needs >= 2 GPUs')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert len({r[3] for r in res}) == 1 and res[0][3] > 0

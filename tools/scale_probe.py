"""Per-stage timing of one multi-GPU step (run under torchrun)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bench import build_models, synthetic_scene
from vtaco_b200.conv_onet.generation import Generator3D
from vtaco_b200 import dist as vdist
rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))); torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
group = dist.group.WORLD if world > 1 else None
nx = int(os.environ.get('NX', '256'))
net = build_models(dev)
gen = Generator3D(net, device=dev, resolution0=nx // 4, with_img=True, padding=0.1, input_type='pointcloud')
cloud, tips, tf, touch = synthetic_scene(0)
c = {'grid': torch.randn(1, 64, 64, 64, 32, device=dev).permute(0, 4, 1, 2, 3)}
tips_arg = (tips, torch.from_numpy(tf).to(dev), touch, 0.05)
dec = net.decoder
res = {}
for ex in (['fused', 'nccl'] if world > 1 else ['single']):
    for _ in range(3):
        g, k = gen.eval_lattice(c, tips=tips_arg, group=group, exchange=ex if world > 1 else None)
        gen.mc(g, level_keys=k, sync=False)
    # whole step
    ts = []
    for _ in range(10):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record()
        g, k = gen.eval_lattice(c, tips=tips_arg, group=group, exchange=ex if world > 1 else None)
        e1.record()
        gen.mc(g, level_keys=k, sync=False)
        e2.record()
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    res[ex] = {'lattice_ms': float(np.median([t[0] for t in ts])), 'mc_ms': float(np.median([t[1] for t in ts]))}
    # decode kernel alone on this rank's slab (no exchange)
    x0, x1 = vdist.slab(nx, rank, world)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            dec.forward_dense(c, nx, x0=x0, x1=x1, use_img=True, tips=tips_arg, out=gen._grid if gen._grid is not None else None,
                              axis=gen._axis)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    res[ex]['slab_kernel_local_ms'] = float(np.median(ts))
    if ex == 'fused':
        exo = gen._fused
        ts = []
        for _ in range(10):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            with torch.no_grad():
                dec.forward_dense(c, nx, x0=x0, x1=x1, use_img=True, tips=tips_arg, out=exo.grid, axis=gen._axis, peers=exo.grid_ptrs)
            e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[ex]['slab_kernel_peer_stores_ms'] = float(np.median(ts))
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); exo.barrier(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[ex]['symm_barrier_ms'] = float(np.median(ts))
if rank == 0:
    print(json.dumps({'nx': nx, 'world': world, **res}, indent=1))
if world > 1:
    dist.destroy_process_group()

"""Per-stage device times of one sharded extraction step (exchange='mesh'), eager launches with CUDA events between
the stages, every rank.  torchrun --nproc-per-node N tools/r02_stage_probe.py  (N=1 works too: plain stages)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import build_models, synthetic_scene
from vtaco_b200.conv_onet.generation import Generator3D
from vtaco_b200 import dist as vdist

rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
torch.cuda.set_device(dev)
group = None
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
    group = dist.group.WORLD
nx = 256
net = build_models(dev)
gen = Generator3D(net, device=dev, resolution0=nx // 4, with_img=True, padding=0.1, input_type='pointcloud')
cloud, tips, tf, touch = synthetic_scene(0)
with torch.no_grad():
    c = net.encode_inputs(torch.from_numpy(cloud)[None].to(dev))
    if world > 1:
        flat = c['grid'].permute(0, 2, 3, 4, 1).contiguous()
        dist.broadcast(flat, 0)
        c = {'grid': flat.permute(0, 4, 1, 2, 3)}
tips_arg = (tips, torch.from_numpy(tf).to(dev), touch, 0.05)
res = {}
if world > 1:
    step = lambda: gen.sharded_mesh(c, tips=tips_arg, group=group)   # noqa: E731
    for _ in range(3):
        step()
    gen._settle_sharded(step, group)
    ex = gen._mesh_ex
    x0, x1 = vdist.slab(nx, rank, world)
    xh = min(x1 + 2, nx)
    dec = net.decoder
    T = []
    for it in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        dist.barrier(); torch.cuda.synchronize()
        with torch.no_grad():
            ev[0].record()
            dec.forward_dense(c, nx, x0=x0, x1=xh, use_img=True, tips=tips_arg, out=gen._grid, minmax_key=gen._keys, axis=gen._axis)
            ev[1].record()
            ex.level(gen._keys)
            ev[2].record()
            v, f, counts = gen.mc(gen._grid[x0:xh], level_ptr=ex.level_ptr, x_emit=x1 - x0, x_origin=x0,
                                  voffset=np.float32(nx / 2), vscale=np.float32(1.1 / nx), sync=False)
            ev[3].record()
            ex.push(counts, v, f)
            ev[4].record()
        torch.cuda.synchronize()
        T.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    res = {'rank': rank, 'rows': [x0, xh], 'decode/level/mc/mesh_exchange ms (median of 6)': [round(float(x), 4) for x in np.median(np.array(T), 0)]}
else:
    for _ in range(3):
        grid, keys = gen.eval_lattice(c, tips=tips_arg)
        gen.mc(grid, level_keys=keys, sync=False)
    T = []
    for it in range(6):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record()
        grid, keys = gen.eval_lattice(c, tips=tips_arg)
        ev[1].record()
        gen.mc(grid, level_keys=keys, voffset=np.float32(nx / 2), vscale=np.float32(1.1 / nx), sync=False)
        ev[2].record()
        torch.cuda.synchronize()
        T.append([ev[i].elapsed_time(ev[i + 1]) for i in range(2)])
    res = {'rank': 0, 'decode/mc ms (median of 6, eager)': [round(float(x), 4) for x in np.median(np.array(T), 0)]}
print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    os._exit(0)

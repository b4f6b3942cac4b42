"""marching cubes on a decoded 256^3 lattice: CUDA-event time of phase 1 (classify + scan, nothing emitted) vs the
whole extraction, warm L2 and after an L2 flush."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import build_models, synthetic_scene
from vtaco_b200 import _abi
from vtaco_b200.conv_onet.generation import Generator3D
dev = torch.device('cuda')
net = build_models(dev)
cloud, tips, tf, touch = synthetic_scene(0)
gen = Generator3D(net, device=dev, resolution0=64, with_img=True, padding=0.1, input_type='pointcloud')
with torch.no_grad():
    c = net.encode_inputs(torch.from_numpy(cloud)[None].to(dev))
    grid, keys = gen.eval_lattice(c, tips=(tips, torch.from_numpy(tf).to(dev), touch, 0.05))
    v, f = gen.extract_mesh(grid, keys)
mc = gen.mc
L = _abi.lib()
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)


def run(phase, cold):
    a = _abi.McArgs()
    a.grid, a.nx, a.ny, a.nz = grid.data_ptr(), 256, 256, 256
    a.level_keys, a.n_level_keys = keys.data_ptr(), 1
    a.scratch, a.scratch_bytes = mc._scratch.data_ptr(), mc._scratch.numel()
    a.counts = mc._counts.data_ptr()
    a.vertices, a.vertex_capacity = mc._verts.data_ptr(), mc._verts.size(0)
    a.faces, a.face_capacity = mc._faces.data_ptr(), mc._faces.size(0)
    a.voffset, a.vscale, a.phase = 128.0, 1.1 / 256, phase
    ts = []
    for _ in range(8):
        if cold:
            flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.check(L.vtaco_marching_cubes(C.byref(a), _abi.stream_ptr(dev)), 'mc')
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for cold in (False, True):
    print('cold' if cold else 'warm', 'phase1 %.4f ms' % run(1, cold), 'phase3 %.4f ms' % run(3, cold), mc._counts[:3].tolist())

"""Time marching cubes alone on the decoded 256^3 lattice of the bench scene (L2 flushed between runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_models, synthetic_scene
from vtaco_b200.conv_onet.generation import Generator3D
dev = torch.device('cuda')
net = build_models(dev)
cloud, tips, tf, touch = synthetic_scene(0)
gen = Generator3D(net, device=dev, resolution0=64, with_img=True, padding=0.1, input_type='pointcloud')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    c = net.encode_inputs(torch.from_numpy(cloud)[None].to(dev))
    grid, keys = gen.eval_lattice(c, tips=(tips, torch.from_numpy(tf).to(dev), touch, 0.05))
    for _ in range(3):
        v, f = gen.extract_mesh(grid, keys)
    for cold in (True, False):
        ts = []
        for _ in range(10):
            if cold:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            v, f = gen.extract_mesh(grid, keys)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print('marching cubes 256^3 %s: %.4f ms (min %.4f)  V=%d F=%d' % ('cold' if cold else 'warm', sum(ts) / len(ts), min(ts), v.shape[0], f.shape[0]))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import build_models, synthetic_scene
from vtaco_b200.conv_onet.generation import Generator3D
dev = torch.device('cuda')
net = build_models(dev)
cloud, tips, tf, touch = synthetic_scene(0)
gen = Generator3D(net, device=dev, resolution0=64, with_img=True, padding=0.1, input_type='pointcloud')
with torch.no_grad():
    c = net.encode_inputs(torch.from_numpy(cloud)[None].to(dev))
    for _ in range(3):
        grid, keys = gen.eval_lattice(c, tips=(tips, torch.from_numpy(tf).to(dev), touch, 0.05))
        v, f = gen.extract_mesh(grid, keys)
torch.cuda.synchronize()
print(v.shape, f.shape)

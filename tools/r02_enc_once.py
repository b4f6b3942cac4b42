"""One PointNet forward of the bench scene (for ncu captures of the encoder kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_models, synthetic_scene
dev = torch.device('cuda')
net = build_models(dev)
cloud = torch.from_numpy(synthetic_scene(0)[0])[None].to(dev)
with torch.no_grad():
    for _ in range(3):
        f = net.encoder.pointnet_features(cloud)
torch.cuda.synchronize()
print({k: tuple(v.shape) for k, v in f.items()})

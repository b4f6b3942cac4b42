"""three fused UNet3D forwards (64^3 x 32, VTacO_YCB kwargs) — for `ncu -k regex:conv3d_tc|maxpool2|channel_stats`."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_models, synthetic_scene
dev = torch.device('cuda')
net = build_models(dev)
cloud = torch.from_numpy(synthetic_scene(0)[0])[None].to(dev)
with torch.no_grad():
    fea = net.encoder.pointnet_features(cloud)['grid']
    for _ in range(3):
        out = net.encoder.unet3d(fea)
torch.cuda.synchronize()
print(out.shape)

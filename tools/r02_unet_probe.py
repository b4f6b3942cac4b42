"""UNet3D (VTacO_YCB kwargs, 64^3 x 32) forward: our fused tcgen05 path vs the torch.nn / cuDNN modules
(TF32 on = torch default, and off), CUDA-event times; and the whole encoder.  gpurun_out/r02_unet_probe.json"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import build_models, synthetic_scene


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


dev = torch.device('cuda')
net = build_models(dev)
enc = net.encoder
cloud = torch.from_numpy(synthetic_scene(0)[0])[None].to(dev)
res = {}
with torch.no_grad():
    fea = enc.pointnet_features(cloud)['grid']
    u = enc.unet3d
    res['pointnet_ms'] = timed(lambda: enc.pointnet_features(cloud))
    u.fused = True
    res['unet3d_fused_ms'] = timed(lambda: u(fea))
    g = torch.cuda.CUDAGraph()
    u(fea); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        out = u(fea)
    res['unet3d_fused_graph_ms'] = timed(lambda: g.replay())
    res['encoder_fused_ms'] = timed(lambda: net.encode_inputs(cloud))
    a = u(fea).clone()
    u.fused = False
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        res['unet3d_cudnn_tf32_%s_ms' % tf32] = timed(lambda: u(fea))
        b = u(fea)
        res['fused_vs_cudnn_tf32_%s_maxrel' % tf32] = float((a - b).abs().max() / b.abs().max())
    torch.backends.cudnn.allow_tf32 = True
    res['encoder_cudnn_ms'] = timed(lambda: net.encode_inputs(cloud))
print(json.dumps(res, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/r02_unet_probe.json', 'w'), indent=1)

"""PointNet part of LocalPoolPointnet (grid 64^3, max pooling) at T = 3 640, 10^5, 10^6 points: ms per forward."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from vtaco_b200.encoder import encoder_dict
torch.manual_seed(0)
enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, hidden_dim=32, plane_type='grid', grid_resolution=64).cuda().eval()
with torch.no_grad():
    for b in enc.blocks:
        b.fc_1.weight.normal_(0, 0.1)
res = {}
with torch.no_grad():
    for T in (3640, 100000, 1000000):
        p = (torch.rand(1, T, 3, device='cuda') - 0.5)
        for _ in range(3):
            enc.pointnet_features(p)
        ts = []
        for _ in range(10):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); enc.pointnet_features(p); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res['T%d' % T] = {'ms': float(np.median(ts)), 'Mpts_per_s': T / float(np.median(ts)) / 1e3,
                          'tflops_fp32': T * 53600 / float(np.median(ts)) / 1e9}
print(json.dumps(res, indent=1))
json.dump(res, open('gpurun_out/r02_enc_sweep.json', 'w'), indent=1)

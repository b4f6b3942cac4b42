"""BASELINE.json configs[1], [2], [4]: shape sweep of the hot-path kernels on one B200
(device-resident inputs, CUDA-event timing, median of 5 after 2 warm-ups) with the reference CPU
path (oracle port) timed on bounded sizes.  Writes gpurun_out/sweep.json.

Lives under tests/ (not tools/) because it times the oracle as the CPU baseline, and only tests/,
smoke() and bench.py's baseline legs may execute oracle/.  Not collected by pytest; run it as
`python tests/perf_sweep.py` on the GPU box."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from vtaco_b200.encoder import encoder_dict
from vtaco_b200.conv_onet.models import decoder_dict
from vtaco_b200.mcubes import MarchingCubes
from oracle import convonet as oc

dev = torch.device('cuda')
HBM = 6555.2


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def randomise(m):
    with torch.no_grad():
        for b in m.blocks:
            b.fc_1.weight.normal_(0, 0.1)


torch.manual_seed(0)
dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32, with_contact=True).to(dev).eval(); randomise(dec)
res = {'device': torch.cuda.get_device_name(0), 'decoder_flat': [], 'training_shape': {}, 'encoder_pointnet': [],
       'marching_cubes': [], 'cpu_reference': {}}
grid1 = {'grid': torch.randn(1, 32, 64, 64, 64, device=dev)}
tri1 = {k: torch.randn(1, 32, 64, 64, device=dev) for k in ('xz', 'xy', 'yz')}

with torch.no_grad():
    # ---- decoder, flat random queries, N = 1e5 .. 1e9 (config 5) ----
    for N in (10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8, 10 ** 9):
        p = (torch.rand(1, N, 3, device=dev) - 0.5) * 1.1
        for name, feats in (('grid64', grid1), ('triplane64', tri1)):
            if N == 10 ** 9 and name != 'grid64':
                continue
            ms = timed(lambda: dec(p, feats), reps=3 if N >= 10 ** 8 else 5)
            res['decoder_flat'].append({'features': name, 'N': N, 'ms': ms, 'Gpts_per_s': N / ms / 1e6,
                                        'kernel': 'tcgen05, four tiles per SM (variant 7; forward)'})
        if N == 10 ** 6:
            ci = torch.randn(1, N, 32, device=dev)
            ms = timed(lambda: dec.forward_img(p, grid1, ci))
            res['decoder_flat'].append({'features': 'grid64', 'N': N, 'ms': ms, 'Gpts_per_s': N / ms / 1e6,
                                        'kernel': 'SIMT FFMA2 (forward_img, dense c_img tensor)'})
            ms = timed(lambda: dec.forward_contact(p, grid1))
            res['decoder_flat'].append({'features': 'grid64', 'N': N, 'ms': ms, 'Gpts_per_s': N / ms / 1e6,
                                        'kernel': 'tcgen05, four tiles per SM (variant 7; forward_contact)'})
        del p
        torch.cuda.empty_cache()

    # ---- config 2: training shape, B=32 x 2048 queries, T=3000 (tri-plane encoder + UNet, and grid encoder + UNet3D)
    B, T, N = 32, 3000, 2048
    cloud = (torch.rand(B, T, 3, device=dev) - 0.5)
    q = (torch.rand(B, N, 3, device=dev) - 0.5) * 1.1
    ci = torch.randn(B, N, 32, device=dev)
    enc_t = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, plane_type=['xz', 'xy', 'yz'],
                                                plane_resolution=32, unet=True,
                                                unet_kwargs=dict(depth=4, merge_mode='concat', start_filts=32)).to(dev).eval()
    randomise(enc_t)
    ct = enc_t(cloud)
    res['training_shape']['triplane32_unet'] = {
        'encoder_ms': timed(lambda: enc_t(cloud)), 'encoder_pointnet_part_ms': timed(lambda: enc_t.pointnet_features(cloud)),
        'decoder_forward_ms': timed(lambda: dec(q, ct)), 'decoder_forward_img_ms': timed(lambda: dec.forward_img(q, ct, ci)),
        'B': B, 'T': T, 'N': N}
    enc_g = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, plane_type='grid',
                                                grid_resolution=64, unet3d=True,
                                                unet3d_kwargs=dict(num_levels=4, f_maps=32, in_channels=32,
                                                                   out_channels=32)).to(dev).eval()
    randomise(enc_g)
    B2 = 8
    cg = enc_g(cloud[:B2])
    res['training_shape']['grid64_unet3d'] = {
        'encoder_ms': timed(lambda: enc_g(cloud[:B2])), 'encoder_pointnet_part_ms': timed(lambda: enc_g.pointnet_features(cloud[:B2])),
        'decoder_forward_ms': timed(lambda: dec(q[:B2], cg)), 'decoder_forward_img_ms': timed(lambda: dec.forward_img(q[:B2], cg, ci[:B2])),
        'B': B2, 'T': T, 'N': N}
    del cg, ct
    torch.cuda.empty_cache()

    # ---- encoder PointNet part, T = 1e3 .. 1e6 (config 5) ----
    enc_pg = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, plane_type='grid',
                                                 grid_resolution=64).to(dev).eval(); randomise(enc_pg)
    enc_pt = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32,
                                                 plane_type=['xz', 'xy', 'yz'], plane_resolution=64).to(dev).eval(); randomise(enc_pt)
    for T in (10 ** 3, 10 ** 4, 10 ** 5, 10 ** 6):
        cl = (torch.rand(1, T, 3, device=dev) - 0.5)
        for name, enc, out_bytes in (('grid64', enc_pg, 64 ** 3 * 128), ('triplane64', enc_pt, 3 * 64 * 64 * 128)):
            ms = timed(lambda: enc.pointnet_features(cl))
            alg = T * 140 + out_bytes
            res['encoder_pointnet'].append({'features': name, 'T': T, 'ms': ms, 'Mpts_per_s': T / ms / 1e3,
                                            'algorithmic_GBps': alg / ms / 1e6, 'frac_of_measured_hbm': alg / ms / 1e6 / HBM})

    # ---- marching cubes ----
    mc = MarchingCubes(dev)
    for nx in (128, 256, 512):
        ax = torch.linspace(-1, 1, nx, device=dev)
        x, y, z = torch.meshgrid(ax, ax, ax, indexing='ij')
        vol = (0.7 - torch.sqrt(x * x + y * y + z * z) + 0.05 * torch.sin(9 * x) * torch.cos(7 * y)).contiguous()
        v, f = mc(vol, 0.0)
        V, F = v.shape[0], f.shape[0]
        ms = timed(lambda: mc(vol, 0.0, sync=False))
        alg = 4 * nx ** 3 + 12 * V + 12 * F
        res['marching_cubes'].append({'nx': nx, 'V': V, 'F': F, 'ms': ms, 'algorithmic_GBps': alg / ms / 1e6,
                                      'frac_of_measured_hbm': alg / ms / 1e6 / HBM})
        del vol, x, y, z

# ---- reference CPU path (oracle port), bounded ----
torch.set_num_threads(os.cpu_count() or 1)
W = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
gcpu = {'grid': grid1['grid'].cpu()}
pc = (torch.rand(1, 10 ** 5, 3) - 0.5) * 1.1
with torch.no_grad():
    oc.decoder_forward(pc, gcpu, W)
    t0 = time.perf_counter(); oc.decoder_forward(pc, gcpu, W); dt = time.perf_counter() - t0
res['cpu_reference']['decoder_forward_grid64_N1e5'] = {'s': dt, 'Mpts_per_s': 0.1 / dt, 'threads': torch.get_num_threads()}
We = {k: v.detach().cpu() for k, v in enc_pg.state_dict().items()}
for T in (10 ** 3, 10 ** 4, 10 ** 5):
    cl = torch.rand(1, T, 3) - 0.5
    with torch.no_grad():
        t0 = time.perf_counter(); oc.encoder_pointnet(cl, We, plane_type='grid', reso_grid=64); dt = time.perf_counter() - t0
    res['cpu_reference']['encoder_pointnet_grid64_T%d' % T] = {'s': dt, 'Mpts_per_s': T / dt / 1e6}
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/sweep.json', 'w'), indent=1)
print(json.dumps(res, indent=1))

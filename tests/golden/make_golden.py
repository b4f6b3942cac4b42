#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (jeffsonyu/VTacO,
mounted read-only at /root/reference) on CPU in the build container.

The reference is Python and cannot travel to the GPU box, so its outputs are
frozen here as small fixtures.  Run:  python tests/golden/make_golden.py

What is shimmed (absent third-party packages, no network):
  torch_scatter  -> oracle.convonet.scatter_mean / scatter_max (published 2.0.9 semantics)
  pykdtree, pybullet, trimesh, igl, plyfile, tensorboardX, skimage.measure,
  matplotlib, mpl_toolkits, chumpy-free: only imported, never called on this path.
  ./data/VTacO_mesh/depth_origin.txt — read at import time by
  src/conv_onet/{generation,training,inferencing}.py:17-18.
Nothing from the reference tree is copied; only its numerical outputs are saved.
"""
import os
import sys
import types
import tempfile
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('VTACO_REF', '/root/reference')
sys.path.insert(0, ROOT)


def install_shims():
    from oracle import convonet as oc

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def _scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        return oc.scatter_mean(src, index, dim_size=dim_size, out=out)

    def _scatter_max(src, index, dim=-1, out=None, dim_size=None):
        return oc.scatter_max(src, index, dim_size), None

    mod('torch_scatter', scatter_mean=_scatter_mean, scatter_max=_scatter_max)
    k = mod('pykdtree')
    try:   # same query API (tree.query(x, k) -> (dist, idx)); exact nearest neighbours either way
        from scipy.spatial import cKDTree as _KD
    except Exception:
        _KD = object
    k.kdtree = mod('pykdtree.kdtree', KDTree=_KD)
    mod('pybullet')
    mod('trimesh', Trimesh=object)
    mod('igl')
    mod('plyfile', PlyData=object, PlyElement=object)
    mod('tensorboardX', SummaryWriter=object)
    sk = mod('skimage')
    sk.measure = mod('skimage.measure', marching_cubes=None, block_reduce=None)
    mpl = mod('matplotlib', use=lambda *a, **k: None)
    mpl.pyplot = mod('matplotlib.pyplot')
    mt = mod('mpl_toolkits')
    mt.mplot3d = mod('mpl_toolkits.mplot3d', Axes3D=object)
    for name in ('chumpy',):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name)


def import_reference():
    scratch = tempfile.mkdtemp(prefix='vtaco_ref_')
    os.makedirs(os.path.join(scratch, 'data', 'VTacO_mesh'))
    np.savetxt(os.path.join(scratch, 'data', 'VTacO_mesh', 'depth_origin.txt'), np.zeros(240 * 320))
    os.chdir(scratch)
    install_shims()
    sys.path.insert(0, REF)
    from src import common, layers  # noqa
    from src.encoder import encoder_dict
    from src.conv_onet import models, generation
    return common, encoder_dict, models, generation


def rs_randn(seed, *shape, scale=1.0):
    """Stable cross-version generator (numpy legacy RandomState)."""
    return (np.random.RandomState(seed).randn(*shape) * scale).astype(np.float32)


def rs_uniform(seed, lo, hi, *shape):
    return np.random.RandomState(seed).uniform(lo, hi, size=shape).astype(np.float32)


def randomise(module, seed):
    """Default init, then every parameter re-drawn from a stable stream so the
    fixture does not depend on torch's RNG; ResnetBlockFC.fc_1.weight (zero-init
    in src/layers.py:39) becomes N(0, 0.1^2) so the residual branch is exercised."""
    rs = np.random.RandomState(seed)
    with torch.no_grad():
        for name, prm in sorted(module.named_parameters()):
            fan_in = prm.shape[1] if prm.dim() > 1 else prm.shape[0]
            bound = 1.0 / np.sqrt(max(fan_in, 1))
            if name.endswith('fc_1.weight'):
                val = rs.randn(*prm.shape) * 0.1
            else:
                val = rs.uniform(-bound, bound, size=tuple(prm.shape))
            prm.copy_(torch.from_numpy(val.astype(np.float32)))


def sd_np(module, prefix=''):
    return {prefix + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def synthetic_cloud(seed, n_visual, n_tactile_per_tip=128, n_tips=5):
    """visual points uniform in [-.5,.5]^3 + 5 tactile Gaussian blobs
    (sigma .01) + N(0,.005^2) noise."""
    rs = np.random.RandomState(seed)
    vis = rs.uniform(-0.5, 0.5, size=(n_visual, 3))
    tips = rs.uniform(-0.35, 0.35, size=(n_tips, 3))
    tac = (tips[:, None, :] + rs.randn(n_tips, n_tactile_per_tip, 3) * 0.01).reshape(-1, 3)
    pts = np.concatenate([vis, tac], 0) + rs.randn(n_visual + n_tips * n_tactile_per_tip, 3) * 0.005
    return pts.astype(np.float32), tips.astype(np.float32)


def main():
    torch.set_num_threads(4)
    common, encoder_dict, models, generation = import_reference()
    out = {}

    # ---- G1: coordinate helpers (src/common.py:268-348) ------------------- #
    pts = rs_uniform(1, -0.62, 0.62, 1, 4096, 3)
    edge = np.array([[0.55, -0.55, 0.0], [0.5500055, 0.55055, -0.5500055], [0.7, -0.7, 0.55],
                     [0.549999, -0.549999, 0.275], [1e-8, -1e-8, 0.0], [np.float32(0.55) - 1e-7, 0.3, -0.3]],
                    dtype=np.float32)
    pts[0, :edge.shape[0]] = edge
    g = {'p': pts}
    tp = torch.from_numpy(pts)
    for plane in ('xz', 'xy', 'yz'):
        xy = common.normalize_coordinate(tp.clone(), padding=0.1, plane=plane)
        g['norm_' + plane] = xy.numpy()
        for R in (32, 64, 128):
            g['idx_%s_%d' % (plane, R)] = common.coordinate2index(xy, R).numpy()
    pn = common.normalize_3d_coordinate(tp.clone(), padding=0.1)
    g['norm_grid'] = pn.numpy()
    for R in (32, 64, 128):
        g['idx_grid_%d' % R] = common.coordinate2index(pn, R, coord_type='3d').numpy()
    for nx in (8, 32, 128, 256):
        g['axis_%d' % nx] = (1.1 * common.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx, 1, 1)))[:, 0].numpy()
    g['grid3_4'] = common.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (4, 3, 2)).numpy()
    np.savez_compressed(os.path.join(HERE, 'coords.npz'), **g)

    # ---- G2: LocalDecoder (decoder.py:9-161), grid + tri-plane + mixed ----- #
    for leaky, contact, tag in ((False, True, 'relu'), (True, False, 'leaky')):
        dec = models.decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, with_contact=contact,
                                                  sample_mode='bilinear', hidden_size=32, leaky=leaky)
        randomise(dec, 11 if not leaky else 12)
        dec.eval()
        B, N, Rg, Rp = 2, 1536, 16, 32
        p = rs_uniform(21, -0.6, 0.6, B, N, 3)
        p[0, :edge.shape[0]] = edge
        feats = {'grid': rs_randn(31, B, 32, Rg, Rg, Rg), 'xz': rs_randn(32, B, 32, Rp, Rp),
                 'xy': rs_randn(33, B, 32, Rp, Rp), 'yz': rs_randn(34, B, 32, Rp, Rp)}
        c_img = rs_randn(35, B, N, 32)
        c_img[:, ::3] = 0.0
        g = {'p': p, 'c_img': c_img, 'feat_seeds': np.array([31, 32, 33, 34]),
             'feat_shapes': np.array([Rg, Rp])}
        g.update(sd_np(dec, 'w.'))
        tp, tci = torch.from_numpy(p), torch.from_numpy(c_img)
        combos = {'grid': ['grid'], 'tri': ['xz', 'xy', 'yz'], 'all': ['grid', 'xz', 'xy', 'yz'], 'xz': ['xz']}
        with torch.no_grad():
            for cname, keys in combos.items():
                cp = {k: torch.from_numpy(feats[k]) for k in keys}
                for mode in ('bilinear', 'nearest'):
                    dec.sample_mode = mode
                    g['fwd_%s_%s' % (cname, mode)] = dec(tp, cp).numpy()
                    g['img_%s_%s' % (cname, mode)] = dec.forward_img(tp, cp, tci).numpy()
                    if contact:
                        o, oc_ = dec.forward_contact(tp, cp)
                        g['con_%s_%s' % (cname, mode)] = np.stack([o.numpy(), oc_.numpy()])
                dec.sample_mode = 'bilinear'
            g['sample_grid'] = dec.sample_grid_feature(tp, torch.from_numpy(feats['grid'])).numpy()
            g['sample_xz'] = dec.sample_plane_feature(tp, torch.from_numpy(feats['xz']), plane='xz').numpy()
            g['sample_yz'] = dec.sample_plane_feature(tp, torch.from_numpy(feats['yz']), plane='yz').numpy()
        np.savez_compressed(os.path.join(HERE, 'decoder_%s.npz' % tag), **g)

    # ---- G3: LocalPoolPointnet PointNet part (pointnet.py:135-172) --------- #
    for tag, kw in (('grid', dict(plane_type='grid', grid_resolution=32)),
                    ('tri', dict(plane_type=['xz', 'xy', 'yz'], plane_resolution=32)),
                    ('all_mean', dict(plane_type=['xz', 'xy', 'yz', 'grid'], plane_resolution=16,
                                      grid_resolution=16, scatter_type='mean'))):
        enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, **kw)
        randomise(enc, 41)
        enc.eval()
        cloud0, _ = synthetic_cloud(51, 600, 40)
        cloud1, _ = synthetic_cloud(52, 600, 40)
        p = np.stack([cloud0, cloud1])
        p[1, :edge.shape[0]] = edge
        g = {'p': p}
        g.update(sd_np(enc, 'w.'))
        with torch.no_grad():
            fea = enc(torch.from_numpy(p))
        for k, v in fea.items():
            v = v.numpy()
            flat = v.reshape(v.shape[0], v.shape[1], -1)
            occ = np.abs(flat).sum(1) != 0
            b_idx, cell = np.nonzero(occ)
            g['fea_%s_shape' % k] = np.array(v.shape)
            g['fea_%s_b' % k] = b_idx.astype(np.int32)
            g['fea_%s_cell' % k] = cell.astype(np.int32)
            g['fea_%s_val' % k] = flat[b_idx, :, cell]
        g['key_order'] = np.array(list(fea.keys()))
        np.savez_compressed(os.path.join(HERE, 'encoder_%s.npz' % tag), **g)

    # ---- G4: Generator3D.eval_points on the dense lattice (generation.py:338-383)
    dec = models.decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, with_contact=False,
                                              sample_mode='bilinear', hidden_size=32)
    randomise(dec, 61)
    net = models.ConvolutionalOccupancyNetwork(dec, None, None, None, None, device='cpu')
    nx = 32
    gen = generation.Generator3D(net, device='cpu', resolution0=nx // 4, with_img=True, padding=0.1,
                                 input_type='pointcloud', points_batch_size=10000)
    pointsf = 1.1 * common.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx,) * 3)
    Rg = 16
    c = {'grid': torch.from_numpy(rs_randn(71, 1, 32, Rg, Rg, Rg))}
    tips = rs_uniform(72, -0.3, 0.3, 5, 3)
    tip_feat = rs_randn(73, 5, 32)
    touch = np.array([1, 0, 1, 1, 0], dtype=bool)
    from oracle.convonet import fingertip_c_img
    c_img_all = fingertip_c_img(pointsf, tips, torch.from_numpy(tip_feat), touch, 0.05)
    vals_img = gen.eval_points(pointsf, c, c_img_all.unsqueeze(0)).numpy()
    gen.with_img = False
    vals = gen.eval_points(pointsf, c).numpy()
    g = {'nx': np.array(nx), 'feat_seed': np.array(71), 'Rg': np.array(Rg), 'tips': tips, 'tip_feat': tip_feat,
         'touch': touch, 'logits_img': vals_img, 'logits': vals,
         'c_img_rows': np.nonzero(np.abs(c_img_all.numpy()).sum(1))[0].astype(np.int32)}
    g.update(sd_np(dec, 'w.'))
    np.savez_compressed(os.path.join(HERE, 'eval_points.npz'), **g)

    # ---- G5: state_dict contract + UNet / UNet3D post-processing (unet.py:117-239, unet3d.py:361-491)
    import json
    from src.encoder.unet import UNet
    from src.encoder.unet3d import UNet3D
    contract = {}
    enc_g = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, plane_type='grid',
                                                grid_resolution=64, unet3d=True,
                                                unet3d_kwargs=dict(num_levels=4, f_maps=32, in_channels=32, out_channels=32))
    enc_t = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32,
                                                plane_type=['xz', 'xy', 'yz'], plane_resolution=32, unet=True,
                                                unet_kwargs=dict(depth=4, merge_mode='concat', start_filts=32))
    dec_c = models.decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, with_contact=True,
                                                sample_mode='bilinear', hidden_size=32)
    for name, m in (('encoder_grid_unet3d', enc_g), ('encoder_tri_unet', enc_t), ('decoder_contact', dec_c)):
        contract[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
    with open(os.path.join(HERE, 'state_dict_contract.json'), 'w') as f:
        json.dump(contract, f, indent=0, sort_keys=True)
    g = {}
    u2 = UNet(8, in_channels=8, depth=3, merge_mode='concat', start_filts=8)
    randomise(u2, 81)
    u3 = UNet3D(in_channels=8, out_channels=8, num_levels=3, f_maps=8)
    randomise(u3, 82)
    x2, x3 = rs_randn(83, 2, 8, 16, 16), rs_randn(84, 1, 8, 8, 8, 8)
    with torch.no_grad():
        g['unet_out'] = u2.eval()(torch.from_numpy(x2)).numpy()
        g['unet3d_out'] = u3.eval()(torch.from_numpy(x3)).numpy()
    g.update(sd_np(u2, 'u2.'))
    g.update(sd_np(u3, 'u3.'))
    np.savez_compressed(os.path.join(HERE, 'unets.npz'), **g)

    make_grads(common, encoder_dict, models, generation)
    make_encoder_grads(common, encoder_dict, models, generation)
    make_chamfer(common)
    make_emd(common)
    make_helpers(common)
    make_unet3d_shipped()
    make_unet2d_shipped()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')


def make_grads(common, encoder_dict, models, generation):
    """G6: gradients torch autograd produces through the reference LocalDecoder
    (what training.py:79,617 back-propagates): loss = sum(logits * r) [+ sum(contact * r2)]."""
    B, N, Rg, Rp = 2, 320, 8, 16
    edge = np.array([[0.55, -0.55, 0.0], [0.7, -0.7, 0.55], [0.549999, -0.549999, 0.275]], dtype=np.float32)
    p = rs_uniform(121, -0.6, 0.6, B, N, 3)
    p[0, :edge.shape[0]] = edge
    feats = {'grid': rs_randn(131, B, 32, Rg, Rg, Rg), 'xz': rs_randn(132, B, 32, Rp, Rp),
             'xy': rs_randn(133, B, 32, Rp, Rp), 'yz': rs_randn(134, B, 32, Rp, Rp)}
    c_img = rs_randn(135, B, N, 32)
    c_img[:, ::3] = 0.0
    r, r2 = rs_randn(136, B, N), rs_randn(137, B, N)
    g = {'p': p, 'c_img': c_img, 'r': r, 'r2': r2, 'feat_seeds': np.array([131, 132, 133, 134]),
         'feat_shapes': np.array([Rg, Rp])}
    cases = (('img_grid_relu', False, ['grid'], 'img', 'bilinear'),
             ('fwd_tri_leaky', True, ['xz', 'xy', 'yz'], 'fwd', 'bilinear'),
             ('con_all_relu', False, ['grid', 'xz', 'xy', 'yz'], 'con', 'bilinear'),
             ('img_all_nearest', False, ['grid', 'xz'], 'img', 'nearest'))
    for tag, leaky, keys, mode, smode in cases:
        dec = models.decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, with_contact=(mode == 'con'),
                                                  sample_mode=smode, hidden_size=32, leaky=leaky)
        randomise(dec, 111)
        dec.train()
        cp = {k: torch.from_numpy(feats[k]).requires_grad_(True) for k in keys}
        tci = torch.from_numpy(c_img).requires_grad_(True)
        tp = torch.from_numpy(p)
        if mode == 'img':
            loss = (dec.forward_img(tp, cp, tci) * torch.from_numpy(r)).sum()
        elif mode == 'con':
            o, oc_ = dec.forward_contact(tp, cp)
            loss = (o * torch.from_numpy(r)).sum() + (oc_ * torch.from_numpy(r2)).sum()
        else:
            loss = (dec(tp, cp) * torch.from_numpy(r)).sum()
        loss.backward()
        g[tag + '.loss'] = np.array(loss.item())
        for k, v in sd_np(dec, tag + '.w.').items():
            g[k] = v
        for n, prm in dec.named_parameters():
            if prm.grad is not None:
                g['%s.dw.%s' % (tag, n)] = prm.grad.numpy()
        for k in keys:
            g['%s.dfeat.%s' % (tag, k)] = cp[k].grad.numpy()
        if mode == 'img':
            g[tag + '.dc_img'] = tci.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'decoder_grads.npz'), **g)


ENC_GRAD_CASES = (('grid_max', dict(plane_type='grid', grid_resolution=16)),
                  ('tri_max', dict(plane_type=['xz', 'xy', 'yz'], plane_resolution=16)),
                  ('all_mean', dict(plane_type=['xz', 'xy', 'yz', 'grid'], plane_resolution=8, grid_resolution=8,
                                    scatter_type='mean')))


def make_encoder_grads(common, encoder_dict, models, generation):
    """G7: parameter gradients torch autograd produces through the reference LocalPoolPointnet
    (PointNet part, no UNet): loss = sum_key sum(fea[key] * r_key).  torch_scatter is shimmed by
    the oracle's restatement (see install_shims), everything else is the reference's code."""
    cloud0, _ = synthetic_cloud(151, 260, 20)
    cloud1, _ = synthetic_cloud(152, 260, 20)
    p = np.stack([cloud0, cloud1])
    g = {'p': p}
    for tag, kw in ENC_GRAD_CASES:
        enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, **kw)
        randomise(enc, 141)
        enc.train()
        fea = enc(torch.from_numpy(p))
        loss = 0
        for i, (k, v) in enumerate(fea.items()):
            r = rs_randn(160 + i, *v.shape)
            loss = loss + (v * torch.from_numpy(r)).sum()
        loss.backward()
        g[tag + '.loss'] = np.array(loss.item())
        g[tag + '.keys'] = np.array(list(fea.keys()))
        for k, v in sd_np(enc, tag + '.w.').items():
            g[k] = v
        for n, prm in enc.named_parameters():
            g['%s.dw.%s' % (tag, n)] = prm.grad.numpy()
    np.savez_compressed(os.path.join(HERE, 'encoder_grads.npz'), **g)


def make_chamfer(common):
    """G8: src/common.py chamfer_distance on mesh-vertex-like sets (generation.py:281 uses 2048 points)."""
    g = {}
    for tag, B, T in (('t2048', 2, 2048), ('t300', 3, 300)):
        a = rs_uniform(171, -0.5, 0.5, B, T, 3)
        b = (a + rs_randn(172, B, T, 3, scale=0.02))[:, ::-1].copy()
        g[tag + '.p1'], g[tag + '.p2'] = a, b
        ta, tb = torch.from_numpy(a), torch.from_numpy(b)
        g[tag + '.naive'] = common.chamfer_distance(ta, tb, use_kdtree=False).numpy()
        c1, c2, i12, i21 = common.chamfer_distance(ta, tb, use_kdtree=True, give_id=True)
        g[tag + '.kd_c1'], g[tag + '.kd_c2'] = c1.numpy(), c2.numpy()
        g[tag + '.kd_i12'], g[tag + '.kd_i21'] = i12.numpy().astype(np.int32), i21.numpy().astype(np.int32)
    np.savez_compressed(os.path.join(HERE, 'chamfer.npz'), **g)


def make_emd(common):
    """G9: src/common.py EarthMoverDistance (scipy cdist + linear_sum_assignment), generation.py:282."""
    g = {}
    for tag, n1, n2 in (('t2048', 2048, 2048), ('t300', 300, 300), ('rect', 256, 200), ('rect2', 120, 300)):
        a = rs_uniform(181, -0.5, 0.5, n1, 3)
        b = rs_uniform(182, -0.5, 0.5, n2, 3)
        k = min(n1, n2) // 2
        b[:k] = a[:k][::-1] + rs_randn(183, k, 3, scale=0.01)      # half of the points have a near partner
        g[tag + '.p1'], g[tag + '.p2'] = a, b
        g[tag + '.emd'] = np.float64(common.EarthMoverDistance(a, b))
    np.savez_compressed(os.path.join(HERE, 'emd.npz'), **g)


def make_unet3d_shipped():
    """G11: the reference's UNet3D with the SHIPPED kwargs (VTacO_YCB.yaml:26-31: num_levels 4, f_maps 32,
    32 -> 32 channels) on a 16^3 grid, fp32 on the CPU.  Parameters come from `randomise(module, 91)` (a numpy
    stream), so the test re-creates them instead of storing 4.1 M weights."""
    from src.encoder.unet3d import UNet3D
    u = UNet3D(in_channels=32, out_channels=32, num_levels=4, f_maps=32)
    randomise(u, 91)
    x = rs_randn(92, 1, 32, 16, 16, 16) * (np.random.RandomState(93).rand(1, 32, 16, 16, 16) < 0.2)
    with torch.no_grad():
        y = u.eval()(torch.from_numpy(x.astype(np.float32))).numpy()
    np.savez_compressed(os.path.join(HERE, 'unet3d_shipped.npz'), y=y, seed_w=np.int64(91), seed_x=np.array([92, 93]))


def make_unet2d_shipped():
    """G12: the reference's 2-D UNet with the SHIPPED kwargs (VTacO_YCB.yaml:41-44: depth 4, merge_mode concat,
    start_filts 32; 32 -> 32 channels) on two 32 x 32 feature planes, fp32 on the CPU.  Parameters come from
    `randomise(module, 95)` (a numpy stream), so the test re-creates them instead of storing 7.8 M weights."""
    from src.encoder.unet import UNet
    u = UNet(32, in_channels=32, depth=4, merge_mode='concat', start_filts=32)
    randomise(u, 95)
    x = rs_randn(96, 2, 32, 32, 32) * (np.random.RandomState(97).rand(2, 32, 32, 32) < 0.3)
    with torch.no_grad():
        y = u.eval()(torch.from_numpy(x.astype(np.float32))).numpy()
    np.savez_compressed(os.path.join(HERE, 'unet2d_shipped.npz'), y=y, seed_w=np.int64(95), seed_x=np.array([96, 97]))


def make_helpers(common):
    """G10: src/common.py R_from_PYR / norm_pc_1 and the fingertip transform of generation.py:177-188
    (the inline code there, evaluated with the reference's own helper functions)."""
    rs = np.random.RandomState(191)
    g = {}
    angles = rs.uniform(-np.pi, np.pi, size=(4, 3))
    g['angles'] = angles
    g['R'] = np.stack([common.R_from_PYR(a) for a in angles])
    pc_obj = rs.randn(500, 3) * 0.1 + np.array([0.3, -0.2, 0.5])
    pc = rs.randn(7, 3) * 0.2
    g['pc_obj'], g['pc'] = pc_obj, pc
    g['norm'] = common.norm_pc_1(pc, pc_obj)
    joints = rs.randn(21, 3).astype(np.float32) * 0.05
    wrist_rot, wrist_pos = rs.uniform(-1, 1, 3), rs.randn(3) * 0.1
    tips_pos = joints[[4, 8, 12, 16, 20]]
    tips_pos = tips_pos - np.array([0.11, 0.005, 0], dtype=np.float32)
    tips_pos = np.linalg.inv(common.R_from_PYR(np.array([-np.pi / 2, np.pi / 2, 0]))) @ tips_pos.T
    tips_pos = np.linalg.inv(common.R_from_PYR(np.array(wrist_rot))) @ tips_pos
    tips_pos_b = tips_pos.T + wrist_pos
    g['joints'], g['wrist_rot'], g['wrist_pos'] = joints, wrist_rot, wrist_pos
    g['tips'] = common.norm_pc_1(tips_pos_b, pc_obj)
    np.savez_compressed(os.path.join(HERE, 'helpers.npz'), **g)


if __name__ == '__main__':
    if sys.argv[1:] == ['emd']:
        make_emd(import_reference()[0])
    elif sys.argv[1:] == ['helpers']:
        make_helpers(import_reference()[0])
    elif sys.argv[1:] == ['unet3d']:
        import_reference()
        make_unet3d_shipped()
    elif sys.argv[1:] == ['unet2d']:
        import_reference()
        make_unet2d_shipped()
    elif sys.argv[1:] == ['chamfer']:
        make_chamfer(import_reference()[0])
    elif sys.argv[1:] == ['grads']:
        torch.set_num_threads(4)
        ref = import_reference()
        make_grads(*ref)
        make_encoder_grads(*ref)
    else:
        main()

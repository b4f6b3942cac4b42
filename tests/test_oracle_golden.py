"""CPU: the oracle restatement (oracle/convonet.py) against vectors produced by
the reference itself (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import convonet as oc
from util import load, weights, decoder_feats, COMBOS, close, rs_randn


def test_coords_bit_exact():
    g = load('coords.npz')
    p = torch.from_numpy(g['p'])
    for plane in ('xz', 'xy', 'yz'):
        xy = oc.normalize_coordinate(p.clone(), 0.1, plane)
        assert np.array_equal(xy.numpy(), g['norm_' + plane])
        for R in (32, 64, 128):
            assert np.array_equal(oc.coordinate2index(xy, R).numpy(), g['idx_%s_%d' % (plane, R)])
    pn = oc.normalize_3d_coordinate(p.clone(), 0.1)
    assert np.array_equal(pn.numpy(), g['norm_grid'])
    for R in (32, 64, 128):
        assert np.array_equal(oc.coordinate2index(pn, R, '3d').numpy(), g['idx_grid_%d' % R])
    for nx in (8, 32, 128, 256):
        ax = oc.dense_grid_points(nx).view(nx, nx, nx, 3)[:, 0, 0, 0]
        assert np.array_equal(ax.numpy(), g['axis_%d' % nx])
    assert np.array_equal(oc.make_3d_grid((-0.5,) * 3, (0.5,) * 3, (4, 3, 2)).numpy(), g['grid3_4'])


@pytest.mark.parametrize('tag', ['relu', 'leaky'])
def test_decoder_matches_reference(tag):
    g = load('decoder_%s.npz' % tag)
    W = weights(g)
    feats = decoder_feats(g)
    p, c_img = torch.from_numpy(g['p']), torch.from_numpy(g['c_img'])
    leaky = tag == 'leaky'
    with torch.no_grad():
        for cname, keys in COMBOS.items():
            cp = {k: feats[k] for k in keys}
            for mode in ('bilinear', 'nearest'):
                o = oc.decoder_forward(p, cp, W, 'forward', leaky=leaky, sample_mode=mode)
                assert close(o.numpy(), g['fwd_%s_%s' % (cname, mode)]) < 1e-6
                o = oc.decoder_forward(p, cp, W, 'img', c_img=c_img, leaky=leaky, sample_mode=mode)
                assert close(o.numpy(), g['img_%s_%s' % (cname, mode)]) < 1e-6
                if not leaky:
                    o, c_ = oc.decoder_forward(p, cp, W, 'contact', sample_mode=mode)
                    assert close(np.stack([o.numpy(), c_.numpy()]), g['con_%s_%s' % (cname, mode)]) < 1e-6
        assert close(oc.sample_grid_feature(p, feats['grid']).numpy(), g['sample_grid']) < 1e-6
        assert close(oc.sample_plane_feature(p, feats['xz'], 'xz').numpy(), g['sample_xz']) < 1e-6
        assert close(oc.sample_plane_feature(p, feats['yz'], 'yz').numpy(), g['sample_yz']) < 1e-6


ENC_KW = {'grid': dict(plane_type='grid', reso_grid=32),
          'tri': dict(plane_type=['xz', 'xy', 'yz'], reso_plane=32),
          'all_mean': dict(plane_type=['xz', 'xy', 'yz', 'grid'], reso_plane=16, reso_grid=16,
                           scatter_type='mean')}


@pytest.mark.parametrize('tag', ['grid', 'tri', 'all_mean'])
def test_encoder_matches_reference(tag):
    g = load('encoder_%s.npz' % tag)
    W = weights(g)
    with torch.no_grad():
        fea = oc.encoder_pointnet(torch.from_numpy(g['p']), W, **ENC_KW[tag])
    assert list(fea.keys()) == [str(k) for k in g['key_order']]
    for k, v in fea.items():
        assert tuple(v.shape) == tuple(g['fea_%s_shape' % k])
        flat = v.numpy().reshape(v.shape[0], v.shape[1], -1)
        occ = np.abs(flat).sum(1) != 0
        b_idx, cell = np.nonzero(occ)
        assert np.array_equal(b_idx, g['fea_%s_b' % k]) and np.array_equal(cell, g['fea_%s_cell' % k])
        assert close(flat[b_idx, :, cell], g['fea_%s_val' % k]) < 1e-6


def test_eval_points_matches_reference():
    g = load('eval_points.npz')
    W = weights(g)
    nx, Rg = int(g['nx']), int(g['Rg'])
    c = {'grid': torch.from_numpy(rs_randn(int(g['feat_seed']), 1, 32, Rg, Rg, Rg))}
    p = oc.dense_grid_points(nx)
    c_img_all = oc.fingertip_c_img(p, g['tips'], torch.from_numpy(g['tip_feat']), g['touch'], 0.05)
    assert np.array_equal(np.nonzero(np.abs(c_img_all.numpy()).sum(1))[0], g['c_img_rows'])
    v = oc.eval_points(p, c, W, c_img_all, points_batch_size=10000)
    assert close(v.numpy(), g['logits_img']) < 1e-6
    v = oc.eval_points(p, c, W, None, points_batch_size=10000)
    assert close(v.numpy(), g['logits']) < 1e-6


def test_cuda_division_flips_few_indices():
    """reciprocal-multiply (CUDA ATen) vs true division (CPU ATen)
    differ in the last bit but only flip ~1e-6 of the cell indices."""
    p = torch.from_numpy(np.random.RandomState(5).uniform(-0.55, 0.55, size=(1, 1000000, 3)).astype(np.float32))
    a = oc.coordinate2index(oc.normalize_3d_coordinate(p, 0.1, False), 64, '3d')
    b = oc.coordinate2index(oc.normalize_3d_coordinate(p, 0.1, True), 64, '3d')
    assert (a != b).sum().item() < 50


GRAD_CASES = (('img_grid_relu', False, ['grid'], 'img', 'bilinear'),
              ('fwd_tri_leaky', True, ['xz', 'xy', 'yz'], 'forward', 'bilinear'),
              ('con_all_relu', False, ['grid', 'xz', 'xy', 'yz'], 'contact', 'bilinear'),
              ('img_all_nearest', False, ['grid', 'xz'], 'img', 'nearest'))


def oracle_decoder_grads(g, tag, leaky, keys, mode, smode, device='cpu'):
    """torch autograd through the oracle restatement: the checker for vtaco_decoder_backward."""
    W = {k[len(tag) + 3:]: torch.from_numpy(v).to(device).requires_grad_(True)
         for k, v in g.items() if k.startswith(tag + '.w.')}
    feats = decoder_feats(g, device)
    cp = {k: feats[k].clone().requires_grad_(True) for k in keys}
    p = torch.from_numpy(g['p']).to(device)
    c_img = torch.from_numpy(g['c_img']).to(device).requires_grad_(True)
    r, r2 = torch.from_numpy(g['r']).to(device), torch.from_numpy(g['r2']).to(device)
    if mode == 'contact':
        o, c_ = oc.decoder_forward(p, cp, W, 'contact', leaky=leaky, sample_mode=smode)
        loss = (o * r).sum() + (c_ * r2).sum()
    else:
        o = oc.decoder_forward(p, cp, W, mode, c_img=c_img if mode == 'img' else None, leaky=leaky, sample_mode=smode)
        loss = (o * r).sum()
    loss.backward()
    out = {'dw.' + k: v.grad for k, v in W.items() if v.grad is not None}
    out.update({'dfeat.' + k: v.grad for k, v in cp.items()})
    if mode == 'img':
        out['dc_img'] = c_img.grad
    return loss.item(), out


def rel_fro(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize('case', GRAD_CASES, ids=[c[0] for c in GRAD_CASES])
def test_decoder_gradients_match_reference(case):
    """autograd through the oracle == autograd through the reference modules (decoder_grads.npz)."""
    tag, leaky, keys, mode, smode = case
    g = load('decoder_grads.npz')
    loss, grads = oracle_decoder_grads(g, tag, leaky, keys, mode, smode)
    assert abs(loss - float(g[tag + '.loss'])) <= 1e-4 * max(1.0, abs(float(g[tag + '.loss'])))
    n = 0
    for k, v in grads.items():
        assert rel_fro(v.numpy(), g['%s.%s' % (tag, k)]) < 1e-5, k
        n += 1
    assert n == sum(1 for k in g if k.startswith(tag + '.d'))


ENC_GRAD_CASES = (('grid_max', dict(plane_type='grid', reso_grid=16)),
                  ('tri_max', dict(plane_type=['xz', 'xy', 'yz'], reso_plane=16)),
                  ('all_mean', dict(plane_type=['xz', 'xy', 'yz', 'grid'], reso_plane=8, reso_grid=8,
                                    scatter_type='mean')))


def oracle_encoder_grads(g, tag, kw):
    """torch autograd through the oracle's PointNet restatement: the checker for vtaco_encoder_backward."""
    W = {k[len(tag) + 3:]: torch.from_numpy(v).requires_grad_(True) for k, v in g.items() if k.startswith(tag + '.w.')}
    fea = oc.encoder_pointnet(torch.from_numpy(g['p']), W, **kw)
    loss = 0
    for i, (k, v) in enumerate(fea.items()):
        loss = loss + (v * torch.from_numpy(rs_randn(160 + i, *v.shape))).sum()
    loss.backward()
    return loss.item(), list(fea.keys()), {k: v.grad for k, v in W.items()}


@pytest.mark.parametrize('case', ENC_GRAD_CASES, ids=[c[0] for c in ENC_GRAD_CASES])
def test_encoder_gradients_match_reference(case):
    tag, kw = case
    g = load('encoder_grads.npz')
    loss, keys, grads = oracle_encoder_grads(g, tag, kw)
    assert keys == [str(k) for k in g[tag + '.keys']]
    assert abs(loss - float(g[tag + '.loss'])) <= 1e-4 * max(1.0, abs(float(g[tag + '.loss'])))
    for k, v in grads.items():
        assert rel_fro(v.numpy(), g['%s.dw.%s' % (tag, k)]) < 1e-5, k


@pytest.mark.parametrize('tag', ['t2048', 't300'])
def test_chamfer_matches_reference(tag):
    g = load('chamfer.npz')
    a, b = torch.from_numpy(g[tag + '.p1']), torch.from_numpy(g[tag + '.p2'])
    assert close(oc.chamfer_distance_naive(a, b).numpy(), g[tag + '.naive'], 1e-6) < 1e-6
    c1, c2, i12, i21 = oc.chamfer_distance_kdtree(a, b, give_id=True)
    assert np.array_equal(i12.numpy(), g[tag + '.kd_i12']) and np.array_equal(i21.numpy(), g[tag + '.kd_i21'])
    assert close(c1.numpy(), g[tag + '.kd_c1'], 1e-6) < 1e-6 and close(c2.numpy(), g[tag + '.kd_c2'], 1e-6) < 1e-6


@pytest.mark.parametrize('tag', ['t300', 'rect'])
def test_emd_oracle_matches_reference(tag):
    g = load('emd.npz')
    assert abs(oc.earth_mover_distance(g[tag + '.p1'], g[tag + '.p2']) - float(g[tag + '.emd'])) < 1e-12

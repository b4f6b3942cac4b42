"""Compact tactile conditioning on the GPU against the oracle's restatement of the
reference's host code: fingertip ids (generation.py:190-200), tactile point-cloud map
(generation.py:222-255), and the decoder consuming the byte map == the dense c_img_all path."""
import numpy as np
import pytest
import torch

from util import load, weights, close, rs_randn

pytestmark = pytest.mark.gpu


def _dec(W):
    from vtaco_b200.conv_onet.models import decoder_dict
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, with_contact='fc_out_contact.weight' in W,
                                       hidden_size=32)
    dec.load_state_dict(W)
    dec = dec.cuda().eval()
    dec.division = 'true'
    return dec


def test_fingertip_ids_match_oracle():
    from oracle import convonet as oc
    from vtaco_b200.conv_onet import tactile
    rs = np.random.RandomState(5)
    B, N = 3, 20000
    p = torch.from_numpy(rs.uniform(-0.55, 0.55, size=(B, N, 3)).astype(np.float32))
    tips = rs.uniform(-0.4, 0.4, size=(B, 5, 3))
    touch = rs.rand(B, 5) < 0.7
    feat = torch.from_numpy(rs.randn(B, 5, 32).astype(np.float32))
    ids = tactile.fingertip_ids(p.cuda(), tips, touch, 0.05)
    dense = tactile.c_img_from_ids(ids, feat.cuda()).cpu()
    hits = 0
    for b in range(B):
        ref = oc.fingertip_c_img(p[b], tips[b], feat[b], touch[b], 0.05)
        assert torch.equal(dense[b], ref)
        hits += int((ids[b] > 0).sum())
    assert hits > 50


@pytest.mark.parametrize('nx', [32, 96])
def test_tactile_point_map_matches_oracle(nx):
    from oracle import convonet as oc
    from vtaco_b200.conv_onet import tactile
    rs = np.random.RandomState(6)
    centers = rs.uniform(-0.4, 0.4, size=(5, 3))
    pts = [centers[t] + rs.randn(128 if t != 3 else 17, 3) * 0.01 for t in range(5)]
    pts[1] = pts[0][:40] + 0.004          # overlapping sensors: the later one must win
    touch = [True, True, False, True, True]
    feat = torch.from_numpy(rs.randn(5, 32).astype(np.float32))
    lattice = oc.dense_grid_points(nx)
    ref = oc.tactile_points_c_img(lattice, pts, feat, touch, 0.015)
    m = tactile.tactile_point_map(pts, touch, 0.015, nx=nx, device='cuda')
    got = tactile.c_img_from_ids(m[None], feat[None].cuda())[0].cpu()
    assert torch.equal(got, ref) and int((m > 0).sum()) > 5
    # flat queries (arbitrary points) through the brute-force kernel
    q = torch.from_numpy(rs.uniform(-0.45, 0.45, size=(30000, 3)).astype(np.float32))
    q[:400] = torch.from_numpy((np.concatenate(pts)[:400] + rs.randn(400, 3) * 0.01).astype(np.float32))
    ref = oc.tactile_points_c_img(q, pts, feat, touch, 0.015)
    m2 = tactile.tactile_point_map(pts, touch, 0.015, p=q.cuda())
    assert torch.equal(tactile.c_img_from_ids(m2[None], feat[None].cuda())[0].cpu(), ref) and int((m2 > 0).sum()) > 20


@pytest.mark.parametrize('variant', [7, 5, 2])
def test_decoder_byte_map_equals_dense_c_img(variant):
    """forward_img / forward_dense with the one-byte-per-query map == the reference's dense c_img_all."""
    from oracle import convonet as oc
    from vtaco_b200.conv_onet import tactile
    g = load('decoder_relu.npz')
    W = weights(g)
    dec = _dec(W)
    dec.kernel_variant = variant
    rs = np.random.RandomState(7)
    R, nx = 24, 40
    feats = {'grid': torch.from_numpy(rs_randn(31, 1, 32, R, R, R))}
    c = {'grid': feats['grid'].cuda()}
    centers = rs.uniform(-0.4, 0.4, size=(5, 3))
    pts = [centers[t] + rs.randn(100, 3) * 0.012 for t in range(5)]
    touch = [True, False, True, True, True]
    feat = torch.from_numpy(rs.randn(5, 32).astype(np.float32))
    lattice = oc.dense_grid_points(nx)
    c_all = oc.tactile_points_c_img(lattice, pts, feat, touch, 0.015)
    ref = oc.eval_points(lattice, feats, W, c_all)
    m = tactile.tactile_point_map(pts, touch, 0.015, nx=nx, device='cuda')
    with torch.no_grad():
        got = dec.forward_dense(c, nx, use_img=True, tip_map=(m, feat.cuda()))
        assert close(got.reshape(-1).cpu().numpy(), ref.numpy()) < 1e-4
        flat = dec.forward_img(lattice[None].cuda(), c, tip_ids=(m[None], feat.cuda()))
        assert close(flat.reshape(-1).cpu().numpy(), ref.numpy()) < 1e-4
    assert int((m > 0).sum()) > 20


def test_training_samples_construction():
    """training.py:560-612 on the device: tip points first (<= 512 per finger), features attached,
    the rest drawn from the other points; shapes / invariants (the draws are random)."""
    from vtaco_b200.conv_onet import tactile
    rs = np.random.RandomState(8)
    B, N, S = 2, 30000, 2048
    p = torch.from_numpy(rs.uniform(-0.5, 0.5, size=(B, N, 3)).astype(np.float32)).cuda()
    occ = (torch.rand(B, N, device='cuda') > 0.5).float()
    tips = rs.uniform(-0.3, 0.3, size=(B, 5, 3))
    touch = np.array([[1, 1, 0, 1, 1], [1, 0, 1, 1, 1]], dtype=bool)
    c_img = torch.randn(B, 5, 32, device='cuda', requires_grad=True)
    gen = torch.Generator(device='cuda').manual_seed(0)
    ps, on, ci = tactile.build_training_samples(p, occ, c_img, tips, touch, S, generator=gen)
    assert ps.shape == (B, S, 3) and on.shape == (B, S) and ci.shape == (B, S, 32) and ci.requires_grad
    ids = tactile.fingertip_ids(ps, tips, touch, 0.05)
    for b in range(B):
        k = int((ci[b].abs().sum(1) > 0).sum())
        assert k > 0 and bool((ids[b, :k] > 0).all())                     # the leading rows are fingertip points
        assert torch.equal(ci[b, :k].detach(), c_img[b, (ids[b, :k] - 1).long()].detach())
        assert float(ci[b, k:].abs().sum()) == 0.0
    ci.sum().backward()
    assert c_img.grad is not None and float(c_img.grad.abs().sum()) > 0

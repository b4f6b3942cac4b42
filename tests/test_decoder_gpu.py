"""GPU parity of the fused decoder (through the C ABI) against the golden vectors of
the reference and against the oracle on fresh seeded inputs.
Tolerance (north_star): |a-b| <= 1e-4 * max(1, |b|)."""
import numpy as np
import pytest
import torch

from util import load, weights, decoder_feats, COMBOS, close, rs_randn, rs_uniform

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_decoder(W, leaky=False, contact=None, mode='bilinear', division='true'):
    if contact is None:
        contact = 'fc_out_contact.weight' in W
    from vtaco_b200.conv_onet.models import decoder_dict
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, with_contact=contact,
                                       sample_mode=mode, hidden_size=32, leaky=leaky)
    dec.load_state_dict(W, strict=True)
    dec = dec.cuda().eval()
    dec.division = division
    return dec


@pytest.mark.parametrize('variant', [0, 1, 2, 4, 5, 6, 7])
@pytest.mark.parametrize('tag', ['relu', 'leaky'])
def test_decoder_golden(tag, variant):
    g = load('decoder_%s.npz' % tag)
    W = weights(g)
    leaky = tag == 'leaky'
    dec = make_decoder(W, leaky=leaky, contact=not leaky)
    dec.kernel_variant = variant
    feats = decoder_feats(g, 'cuda')
    p = torch.from_numpy(g['p']).cuda()
    c_img = torch.from_numpy(g['c_img']).cuda()
    with torch.no_grad():
        for cname, keys in COMBOS.items():
            cp = {k: feats[k] for k in keys}
            for mode in ('bilinear', 'nearest'):
                dec.sample_mode = mode
                o = dec(p, cp)
                assert o.shape == (2, 1536)
                assert close(o.cpu().numpy(), g['fwd_%s_%s' % (cname, mode)]) < TOL, (cname, mode)
                o = dec.forward_img(p, cp, c_img)
                assert close(o.cpu().numpy(), g['img_%s_%s' % (cname, mode)]) < TOL, (cname, mode)
                if not leaky:
                    o, oc_ = dec.forward_contact(p, cp)
                    got = np.stack([o.cpu().numpy(), oc_.cpu().numpy()])
                    assert close(got, g['con_%s_%s' % (cname, mode)]) < TOL, (cname, mode)
        dec.sample_mode = 'bilinear'
        assert close(dec.sample_grid_feature(p, feats['grid']).cpu().numpy(), g['sample_grid']) < 1e-5
        assert close(dec.sample_plane_feature(p, feats['xz'], 'xz').cpu().numpy(), g['sample_xz']) < 1e-5
        assert close(dec.sample_plane_feature(p, feats['yz'], 'yz').cpu().numpy(), g['sample_yz']) < 1e-5


@pytest.mark.parametrize('variant', [1, 2, 5, 7])
@pytest.mark.parametrize('B,N', [(1, 1), (1, 511), (3, 513), (2, 100000), (32, 2048)])
def test_decoder_vs_oracle_shapes(B, N, variant):
    """ragged / tiny / training-shape batches against the oracle (CPU)."""
    from oracle import convonet as oc
    g = load('decoder_relu.npz')
    W = weights(g)
    dec = make_decoder(W, contact=True)
    dec.kernel_variant = variant
    Rg = 24
    feats = {'grid': torch.from_numpy(rs_randn(7, B, 32, Rg, Rg, Rg))}
    p = torch.from_numpy(rs_uniform(8, -0.6, 0.6, B, N, 3))
    c_img = torch.from_numpy(rs_randn(9, B, N, 32))
    with torch.no_grad():
        fc = {k: v.cuda() for k, v in feats.items()}
        ref = oc.decoder_forward(p, feats, W, 'img', c_img=c_img)
        got = dec.forward_img(p.cuda(), fc, c_img.cuda())
        assert close(got.cpu().numpy(), ref.numpy()) < TOL
        ref = oc.decoder_forward(p, feats, W, 'contact')
        got = dec.forward_contact(p.cuda(), fc)
        assert close(got[0].cpu().numpy(), ref[0].numpy()) < TOL and close(got[1].cpu().numpy(), ref[1].numpy()) < TOL


def test_decoder_empty_and_errors():
    g = load('decoder_relu.npz')
    dec = make_decoder(weights(g))
    feats = decoder_feats(g, 'cuda')
    with torch.no_grad():
        o = dec(torch.zeros(2, 0, 3, device='cuda'), {'grid': feats['grid']})
    assert o.shape == (2, 0)
    o = dec(torch.zeros(2, 4, 3, device='cuda'), {'grid': feats['grid']})  # grad enabled: autograd node (decoder_bwd.cu)
    assert o.requires_grad
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            dec(torch.zeros(2, 4, 3), {'grid': feats['grid']})  # CPU tensor: no fallback


def test_channels_last_features_zero_copy():
    g = load('decoder_relu.npz')
    dec = make_decoder(weights(g))
    feats = decoder_feats(g, 'cuda')
    p = torch.from_numpy(g['p']).cuda()
    cl = feats['grid'].contiguous(memory_format=torch.channels_last_3d)
    with torch.no_grad():
        a = dec(p, {'grid': feats['grid']})
        b = dec(p, {'grid': cl})
    assert torch.equal(a, b)


@pytest.mark.parametrize('variant', [1, 2, 5, 6, 7])
def test_dense_matches_flat_and_golden(variant):
    """dense-lattice mode == flat mode on the same lattice; both == reference eval_points."""
    from vtaco_b200.common import make_3d_grid
    g = load('eval_points.npz')
    W = weights(g)
    dec = make_decoder(W)
    dec.kernel_variant = variant
    nx, Rg = int(g['nx']), int(g['Rg'])
    c = {'grid': torch.from_numpy(rs_randn(int(g['feat_seed']), 1, 32, Rg, Rg, Rg)).cuda()}
    pts = (1.1 * make_3d_grid((-0.5,) * 3, (0.5,) * 3, (nx,) * 3)).cuda()
    tip_feat = torch.from_numpy(g['tip_feat']).cuda()
    with torch.no_grad():
        flat = dec(pts[None], c)[0]
        key = torch.tensor([2 ** 31 - 1, -2 ** 31], dtype=torch.int32, device='cuda')
        dense = dec.forward_dense(c, nx, minmax_key=key)
        if variant >= 2:   # the tcgen05 kernel's dense mode interpolates separably (bilinear in x,y then z)
            assert close(dense.reshape(-1).cpu().numpy(), flat.cpu().numpy()) < 1e-5
        else:
            assert torch.equal(dense.reshape(-1), flat)
        assert close(dense.reshape(-1).cpu().numpy(), g['logits']) < TOL
        from vtaco_b200 import _abi
        lo, hi = [_abi.lib().vtaco_key_to_float_host(int(k)) for k in key.cpu()]
        assert lo == dense.min().item() and hi == dense.max().item()
        # compact fingertip conditioning == reference's dense c_img_all
        dimg = dec.forward_dense(c, nx, use_img=True, tips=(g['tips'].astype(np.float64), tip_feat, g['touch'], 0.05))
        assert close(dimg.reshape(-1).cpu().numpy(), g['logits_img']) < TOL
        # slabs reproduce the full grid bit-for-bit
        out = torch.zeros(nx, nx, nx, device='cuda')
        for x0, x1 in ((0, 5), (5, 16), (16, 29), (29, 32)):
            dec.forward_dense(c, nx, x0=x0, x1=x1, out=out)
        assert torch.equal(out, dense)


@pytest.mark.parametrize('variant', [7, 5])
@pytest.mark.parametrize('n_blocks,leaky,mode', [(1, False, 'bilinear'), (3, True, 'bilinear'), (5, False, 'nearest')])
def test_tcgen05_variants_odd_configurations_vs_fp32_kernel(variant, n_blocks, leaky, mode):
    """network depths other than the shipped 5, LeakyReLU head, nearest sampling, grid + planes together, an odd
    lattice (nx = 37: partial 2 x 2 x 32 bricks on every axis, odd slab boundaries) and a ragged flat batch:
    the tensor-core kernels against the exact-fp32 SIMT kernel (variant 1, pinned to the oracle above)."""
    from vtaco_b200.conv_onet.models import decoder_dict
    torch.manual_seed(3)
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, hidden_size=32, n_blocks=n_blocks, leaky=leaky,
                                       sample_mode=mode, with_contact=True).cuda().eval()
    with torch.no_grad():
        for b in dec.blocks:
            b.fc_1.weight.normal_(0, 0.1)
    R = 20
    feats = {'grid': torch.randn(2, 32, R, R, R, device='cuda'), 'xz': torch.randn(2, 32, R, R, device='cuda'),
             'yz': torch.randn(2, 32, R, R, device='cuda')}
    p = (torch.rand(2, 1237, 3, device='cuda') - 0.5) * 1.2
    c_img = torch.randn(2, 1237, 32, device='cuda')
    outs = {}
    with torch.no_grad():
        for v in (1, variant):
            dec.kernel_variant = v
            o = dec.forward_img(p, feats, c_img)
            oc_, occ = dec.forward_contact(p, feats)
            one = {k: t[:1] for k, t in feats.items()}
            d = dec.forward_dense({'grid': one['grid']}, 37)
            part = torch.full((37, 37, 37), float('nan'), device='cuda')
            for x0, x1 in ((0, 5), (5, 6), (6, 37)):
                dec.forward_dense({'grid': one['grid']}, 37, x0=x0, x1=x1, out=part)
            assert torch.equal(part, d)
            outs[v] = (o, oc_, occ, d)
    for a, b in zip(outs[variant], outs[1]):
        assert close(a.cpu().numpy(), b.cpu().numpy()) < 1e-5

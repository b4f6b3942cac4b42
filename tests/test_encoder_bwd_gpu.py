"""GPU parity of vtaco_encoder_backward (through LocalPoolPointnet under autograd) against the
parameter gradients torch autograd produced through the REFERENCE encoder
(tests/golden/encoder_grads.npz) and through the oracle at the shipped cloud size, and an
encoder -> UNet3D -> decoder training step end to end.
Tolerance: relative Frobenius error <= 1e-4 per gradient tensor (fp32 atomic sums)."""
import numpy as np
import pytest
import torch

from util import load, rs_randn, synthetic_cloud
from test_oracle_golden import ENC_GRAD_CASES, rel_fro

pytestmark = pytest.mark.gpu
TOL = 1e-4
CTOR = {'grid_max': dict(plane_type='grid', grid_resolution=16),
        'tri_max': dict(plane_type=['xz', 'xy', 'yz'], plane_resolution=16),
        'all_mean': dict(plane_type=['xz', 'xy', 'yz', 'grid'], plane_resolution=8, grid_resolution=8,
                         scatter_type='mean')}


def make_encoder(W, **kw):
    from vtaco_b200.encoder import encoder_dict
    enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, **kw)
    enc.load_state_dict(W, strict=True)
    enc = enc.cuda().train()
    enc.division = 'true'
    return enc


@pytest.mark.parametrize('tag', [c[0] for c in ENC_GRAD_CASES])
def test_encoder_backward_golden(tag):
    g = load('encoder_grads.npz')
    W = {k[len(tag) + 3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(tag + '.w.')}
    enc = make_encoder(W, **CTOR[tag])
    fea = enc(torch.from_numpy(g['p']).cuda())
    assert list(fea.keys()) == [str(k) for k in g[tag + '.keys']]
    loss = 0
    for i, (k, v) in enumerate(fea.items()):
        assert v.requires_grad
        loss = loss + (v * torch.from_numpy(rs_randn(160 + i, *v.shape)).cuda()).sum()
    loss.backward()
    assert abs(loss.item() - float(g[tag + '.loss'])) <= 1e-4 * max(1.0, abs(float(g[tag + '.loss'])))
    for n, prm in enc.named_parameters():
        ref = g['%s.dw.%s' % (tag, n)]
        assert prm.grad is not None and tuple(prm.grad.shape) == ref.shape, n
        assert rel_fro(prm.grad.cpu().numpy(), ref) < TOL, (n, rel_fro(prm.grad.cpu().numpy(), ref))


def test_encoder_backward_vs_oracle_shipped_cloud():
    """B=2 clouds of 3000+640 points, grid-32 + planes-32 keys together, max pooling."""
    from oracle import convonet as oc
    from vtaco_b200.encoder import encoder_dict
    torch.manual_seed(3)
    enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32,
                                              plane_type=['xz', 'yz', 'grid'], plane_resolution=32, grid_resolution=32)
    with torch.no_grad():
        for n, prm in enc.named_parameters():
            if n.endswith('fc_1.weight'):
                prm.normal_(0, 0.1)
    p = torch.from_numpy(np.stack([synthetic_cloud(200 + b, 3000)[0] for b in range(2)]))
    ref = {}
    for dt in (torch.float32, torch.float64):   # the oracle in both precisions
        W = {k: v.detach().clone().to(dt).requires_grad_(True) for k, v in enc.state_dict().items()}
        fea = oc.encoder_pointnet(p.to(dt), W, plane_type=['xz', 'yz', 'grid'], reso_plane=32, reso_grid=32)
        rs = {k: torch.from_numpy(rs_randn(300 + i, *v.shape)) for i, (k, v) in enumerate(fea.items())}
        sum((v * rs[k].to(dt)).sum() for k, v in fea.items()).backward()
        ref[dt] = {k: v.grad.double().numpy() for k, v in W.items()}
    enc = enc.cuda().train()
    enc.division = 'true'
    out = enc(p.cuda())
    assert list(out.keys()) == list(fea.keys())
    sum((v * rs[k].cuda()).sum() for k, v in out.items()).backward()
    for n, prm in enc.named_parameters():
        # max pooling makes the gradient discontinuous where two points of a cell are tied to within
        # rounding: the oracle's own fp32-vs-fp64 disagreement (1.6e-4 on fc_pos.weight for this
        # cloud, one flipped arg-max) is the floor of what any fp32 implementation can match
        floor = rel_fro(ref[torch.float32][n], ref[torch.float64][n])
        err = min(rel_fro(prm.grad.cpu().numpy(), ref[dt][n]) for dt in ref)
        assert err < max(TOL, 3 * floor), (n, err, floor)


def test_train_step_encoder_unet3d_decoder():
    """training.py:600-620 in miniature: cloud -> LocalPoolPointnet(+UNet3D) -> LocalDecoder -> BCE;
    every parameter of the three parts receives a finite gradient and Adam reduces the loss."""
    import torch.nn.functional as F
    from vtaco_b200.encoder import encoder_dict
    from vtaco_b200.conv_onet.models import decoder_dict
    torch.manual_seed(5)
    enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, plane_type='grid',
                                              grid_resolution=16, unet3d=True,
                                              unet3d_kwargs=dict(num_levels=2, f_maps=32, in_channels=32, out_channels=32)).cuda().train()
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, hidden_size=32).cuda().train()
    rs = np.random.RandomState(11)
    dirs = rs.randn(2, 1500, 3)
    cloud = torch.from_numpy((0.3 * dirs / np.linalg.norm(dirs, axis=2, keepdims=True)).astype(np.float32)).cuda()
    q = torch.from_numpy(rs.uniform(-0.5, 0.5, size=(2, 2048, 3)).astype(np.float32)).cuda()
    occ = (q.norm(dim=2) < 0.3).float()
    params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.Adam(params, lr=2e-3)
    losses = []
    for it in range(40):
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(dec(q, enc(cloud)), occ)
        loss.backward()
        if it == 0:
            for n, prm in list(enc.named_parameters()) + list(dec.named_parameters()):
                if n.startswith('fc_p_img'):
                    continue
                assert prm.grad is not None and torch.isfinite(prm.grad).all(), n
            assert enc.fc_pos.weight.grad.abs().sum() > 0
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.7 * losses[0], losses[::8]

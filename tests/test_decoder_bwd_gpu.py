"""GPU parity of vtaco_decoder_backward (through LocalDecoder under autograd) against the
gradients torch autograd produced through the REFERENCE modules (tests/golden/decoder_grads.npz)
and through the oracle on larger seeded inputs.

Tolerance: relative Frobenius error <= 1e-4 per gradient tensor (fp32 sums over up to 10^4
queries accumulate in a different order; a ReLU whose pre-activation is within rounding of 0
may flip for single queries, which a max-norm would over-weight)."""
import numpy as np
import pytest
import torch

from util import load, decoder_feats, rs_randn, rs_uniform
from test_oracle_golden import GRAD_CASES, oracle_decoder_grads, rel_fro
from test_decoder_gpu import make_decoder

pytestmark = pytest.mark.gpu
TOL = 1e-4


def run_case(g, tag, leaky, keys, mode, smode, variant=5, max_queries=None):
    W = {k[len(tag) + 3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith(tag + '.w.')}
    dec = make_decoder(W, leaky=leaky, contact=(mode == 'contact'), mode=smode)
    dec.train()
    dec.kernel_variant = variant
    if max_queries:
        dec._BWD_MAX_QUERIES = max_queries
    feats = decoder_feats(g, 'cuda')
    cp = {k: feats[k].clone().requires_grad_(True) for k in keys}
    p = torch.from_numpy(g['p']).cuda()
    c_img = torch.from_numpy(g['c_img']).cuda().requires_grad_(True)
    r, r2 = torch.from_numpy(g['r']).cuda(), torch.from_numpy(g['r2']).cuda()
    if mode == 'contact':
        o, c_ = dec.forward_contact(p, cp)
        loss = (o * r).sum() + (c_ * r2).sum()
    elif mode == 'img':
        loss = (dec.forward_img(p, cp, c_img) * r).sum()
    else:
        loss = (dec(p, cp) * r).sum()
    loss.backward()
    out = {'dw.' + n: prm.grad for n, prm in dec.named_parameters() if prm.grad is not None}
    out.update({'dfeat.' + k: v.grad for k, v in cp.items()})
    if mode == 'img':
        out['dc_img'] = c_img.grad
    return loss.item(), out


@pytest.mark.parametrize('variant', [1, 5, 7])
@pytest.mark.parametrize('case', GRAD_CASES, ids=[c[0] for c in GRAD_CASES])
def test_backward_golden(case, variant):
    tag, leaky, keys, mode, smode = case
    g = load('decoder_grads.npz')
    loss, grads = run_case(g, tag, leaky, keys, mode, smode, variant)
    assert abs(loss - float(g[tag + '.loss'])) <= 1e-4 * max(1.0, abs(float(g[tag + '.loss'])))
    want = [k[len(tag) + 1:] for k in g if k.startswith(tag + '.d')]
    assert sorted(grads.keys()) == sorted(want)       # unused heads (fc_p vs fc_p_img) get no gradient
    for k in want:
        ref = g['%s.%s' % (tag, k)]
        got = grads[k].cpu().numpy()
        assert got.shape == ref.shape, k
        assert rel_fro(got, ref) < TOL, (k, rel_fro(got, ref))


def test_backward_chunked_equals_single_launch():
    tag, leaky, keys, mode, smode = GRAD_CASES[0]
    g = load('decoder_grads.npz')
    _, a = run_case(g, tag, leaky, keys, mode, smode)
    _, b = run_case(g, tag, leaky, keys, mode, smode, max_queries=100)   # 2 samples x 4 pieces
    for k in a:
        assert rel_fro(b[k].cpu().numpy(), a[k].cpu().numpy()) < 1e-5, k


def test_backward_training_shape_vs_oracle():
    """B=8 x N=2048 queries on 32^3 grid + 32^2 planes (SURVEY §8d config 2/3 shapes, reduced batch)."""
    from oracle import convonet as oc
    from vtaco_b200.conv_onet.models import decoder_dict
    B, N, Rg, Rp = 8, 2048, 32, 32
    torch.manual_seed(0)
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, sample_mode='bilinear', hidden_size=32)
    with torch.no_grad():
        for n, prm in dec.named_parameters():
            if n.endswith('fc_1.weight'):
                prm.normal_(0, 0.1)
    W = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    p = torch.from_numpy(rs_uniform(5, -0.55, 0.55, B, N, 3))
    feats = {'grid': torch.from_numpy(rs_randn(6, B, 32, Rg, Rg, Rg)), 'xz': torch.from_numpy(rs_randn(7, B, 32, Rp, Rp)),
             'yz': torch.from_numpy(rs_randn(8, B, 32, Rp, Rp))}
    c_img = torch.from_numpy(rs_randn(9, B, N, 32))
    r = torch.from_numpy(rs_randn(10, B, N))
    cp = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    ci = c_img.clone().requires_grad_(True)
    (oc.decoder_forward(p, cp, W, 'img', c_img=ci) * r).sum().backward()

    dec = dec.cuda().train()
    dec.division = 'true'
    cpg = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    cig = c_img.cuda().requires_grad_(True)
    (dec.forward_img(p.cuda(), cpg, cig) * r.cuda()).sum().backward()
    for n, prm in dec.named_parameters():
        if n.startswith('fc_p.'):
            assert prm.grad is None
            continue
        assert rel_fro(prm.grad.cpu().numpy(), W[n].grad.numpy()) < TOL, n
    for k in feats:
        assert rel_fro(cpg[k].grad.cpu().numpy(), cp[k].grad.numpy()) < TOL, k
        assert cpg[k].grad.shape == feats[k].shape
    assert rel_fro(cig.grad.cpu().numpy(), ci.grad.numpy()) < TOL
    # linearity in the incoming gradient: 2x dlogits -> 2x every gradient
    g1 = {n: prm.grad.clone() for n, prm in dec.named_parameters() if prm.grad is not None}
    dec.zero_grad()
    (dec.forward_img(p.cuda(), {k: v.detach() for k, v in cpg.items()}, cig.detach()) * (2 * r.cuda())).sum().backward()
    for n, prm in dec.named_parameters():
        if prm.grad is not None:
            assert rel_fro(prm.grad.cpu().numpy(), 2 * g1[n].cpu().numpy()) < 1e-5, n


def test_backward_trains():
    """A few optimiser steps through the kernels reduce a BCE loss (training.py:600-620 in miniature)."""
    import torch.nn.functional as F
    from vtaco_b200.conv_onet.models import decoder_dict
    torch.manual_seed(1)
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, hidden_size=32).cuda().train()
    feat = torch.from_numpy(rs_randn(3, 2, 32, 16, 16, 16)).cuda().requires_grad_(True)
    p = torch.from_numpy(rs_uniform(4, -0.5, 0.5, 2, 4096, 3)).cuda()
    occ = (p.norm(dim=2) < 0.35).float()
    opt = torch.optim.Adam(list(dec.parameters()) + [feat], lr=1e-2)
    losses = []
    for _ in range(60):
        opt.zero_grad()
        loss = F.binary_cross_entropy_with_logits(dec(p, {'grid': feat}), occ)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.5 * losses[0], losses[::10]


def test_query_points_with_requires_grad_get_no_gradient():
    """The reference training loop builds p with requires_grad=True (training.py:310,362,614,729,868)
    and never reads p.grad: the call must work, p.grad stays None, parameters get their gradients."""
    g = load('decoder_grads.npz')
    W = {k[len('img_grid_relu') + 3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('img_grid_relu.w.')}
    dec = make_decoder(W)
    p = torch.zeros(1, 4, 3, device='cuda', requires_grad=True)
    out = dec(p, {'grid': decoder_feats(g, 'cuda')['grid'][:1]})
    out.sum().backward()
    assert p.grad is None and dec.fc_p.weight.grad is not None

"""Shared helpers for the tests (fixtures, stable random streams)."""
import os
import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def rs_randn(seed, *shape, scale=1.0):
    return (np.random.RandomState(seed).randn(*shape) * scale).astype(np.float32)


def rs_uniform(seed, lo, hi, *shape):
    return np.random.RandomState(seed).uniform(lo, hi, size=shape).astype(np.float32)


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def weights(g, device='cpu'):
    return {k[2:]: torch.from_numpy(v).to(device) for k, v in g.items() if k.startswith('w.')}


def decoder_feats(g, device='cpu'):
    s = g['feat_seeds']
    Rg, Rp = [int(x) for x in g['feat_shapes']]
    B = g['p'].shape[0]
    f = {'grid': rs_randn(int(s[0]), B, 32, Rg, Rg, Rg), 'xz': rs_randn(int(s[1]), B, 32, Rp, Rp),
         'xy': rs_randn(int(s[2]), B, 32, Rp, Rp), 'yz': rs_randn(int(s[3]), B, 32, Rp, Rp)}
    return {k: torch.from_numpy(v).to(device) for k, v in f.items()}


COMBOS = {'grid': ['grid'], 'tri': ['xz', 'xy', 'yz'], 'all': ['grid', 'xz', 'xy', 'yz'], 'xz': ['xz']}


def close(a, b, rtol=1e-4):
    """north_star tolerance: |a-b| <= rtol * max(1, |b|)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    return float(err.max()) if err.size else 0.0


def synthetic_cloud(seed, n_visual, n_tactile_per_tip=128, n_tips=5):
    rs = np.random.RandomState(seed)
    vis = rs.uniform(-0.5, 0.5, size=(n_visual, 3))
    tips = rs.uniform(-0.35, 0.35, size=(n_tips, 3))
    tac = (tips[:, None, :] + rs.randn(n_tips, n_tactile_per_tip, 3) * 0.01).reshape(-1, 3)
    pts = np.concatenate([vis, tac], 0) + rs.randn(n_visual + n_tips * n_tactile_per_tip, 3) * 0.005
    return pts.astype(np.float32), tips.astype(np.float32)


def randomise(module, seed):
    """the parameter stream of tests/golden/make_golden.py: every parameter (sorted by name) re-drawn from
    RandomState(seed); ResnetBlockFC.fc_1.weight ~ N(0, 0.1^2), the rest uniform(+-1/sqrt(fan_in))."""
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

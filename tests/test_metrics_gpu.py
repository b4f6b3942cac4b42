"""GPU parity of vtaco_chamfer (reference src/common.py:54-137) against values produced by the
reference itself (tests/golden/chamfer.npz) and the oracle on ragged sizes.
Neighbour indices: bit-exact (ties do not occur in the seeded sets); distances: 1e-6 relative."""
import numpy as np
import pytest
import torch

from util import load, close, rs_uniform

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['t2048', 't300'])
def test_chamfer_golden(tag):
    from vtaco_b200.common import chamfer_distance
    g = load('chamfer.npz')
    a, b = torch.from_numpy(g[tag + '.p1']).cuda(), torch.from_numpy(g[tag + '.p2']).cuda()
    assert close(chamfer_distance(a, b, use_kdtree=False).cpu().numpy(), g[tag + '.naive'], 1e-6) < 1e-6
    c1, c2, i12, i21 = chamfer_distance(a, b, use_kdtree=True, give_id=True)
    assert i12.dtype == torch.int64
    assert np.array_equal(i12.cpu().numpy(), g[tag + '.kd_i12']) and np.array_equal(i21.cpu().numpy(), g[tag + '.kd_i21'])
    assert close(c1.cpu().numpy(), g[tag + '.kd_c1'], 1e-6) < 1e-6 and close(c2.cpu().numpy(), g[tag + '.kd_c2'], 1e-6) < 1e-6
    assert close(chamfer_distance(a, b).cpu().numpy(), g[tag + '.kd_c1'] + g[tag + '.kd_c2'], 1e-6) < 1e-6


@pytest.mark.parametrize('B,T1,T2', [(1, 1, 1), (2, 1500, 3000), (1, 5000, 257)])
def test_chamfer_ragged_vs_oracle(B, T1, T2):
    from oracle import convonet as oc
    from vtaco_b200.common import chamfer_distance_kdtree
    a, b = torch.from_numpy(rs_uniform(1, -0.5, 0.5, B, T1, 3)), torch.from_numpy(rs_uniform(2, -0.5, 0.5, B, T2, 3))
    r1, r2, j12, j21 = oc.chamfer_distance_kdtree(a, b, give_id=True)
    c1, c2, i12, i21 = chamfer_distance_kdtree(a.cuda(), b.cuda(), give_id=True)
    assert np.array_equal(i12.cpu().numpy(), j12.numpy()) and np.array_equal(i21.cpu().numpy(), j21.numpy())
    assert close(c1.cpu().numpy(), r1.numpy(), 1e-6) < 1e-6 and close(c2.cpu().numpy(), r2.numpy(), 1e-6) < 1e-6


def test_mesh_to_chamfer_end_to_end(tmp_path):
    """sphere logits -> vtaco marching cubes -> OFF file -> chamfer against points on the sphere."""
    from vtaco_b200.mcubes import MarchingCubes
    from vtaco_b200.common import chamfer_distance
    from vtaco_b200.io import export_off, read_off
    nx = 64
    ax = torch.linspace(-0.55, 0.55, nx, device='cuda')
    gx, gy, gz = torch.meshgrid(ax, ax, ax, indexing='ij')
    grid = (0.3 - torch.sqrt(gx * gx + gy * gy + gz * gz)).contiguous()
    mc = MarchingCubes('cuda')
    v, f = mc(grid, level=0.0, voffset=nx / 2, vscale=1.1 / nx)[:2]     # (v - nx/2) * 1.1/nx, generation.py:271-272
    path = str(tmp_path / 'sphere.off')
    export_off(path, v, f)
    v2, f2 = read_off(path)
    assert len(v2) == len(v) and len(f2) == len(f)
    rs = np.random.RandomState(0)
    d = rs.randn(1, 2048, 3)
    gt = torch.from_numpy((0.3 * d / np.linalg.norm(d, axis=2, keepdims=True)).astype(np.float32)).cuda()
    sel = torch.from_numpy(rs.permutation(len(v))[:2048]).cuda()
    cd = chamfer_distance(gt, v[sel][None].contiguous(), use_kdtree=False)
    assert cd.item() < 2 * (0.03 ** 2)      # both sets lie on the same sphere (spacing ~0.02)
    from vtaco_b200.conv_onet.generation import Generator3D
    gen = Generator3D(torch.nn.Identity(), device='cuda', resolution0=16, padding=0.1)
    cd2 = gen.mesh_chamfer(v, gt, generator=torch.Generator(device='cuda').manual_seed(0))
    assert cd2.shape == (1,) and cd2.item() < 2 * (0.03 ** 2)


@pytest.mark.parametrize('tag', ['t300', 'rect', 'rect2', 't2048'])
def test_emd_matches_reference_golden(tag):
    """EarthMoverDistance (auction algorithm on the GPU) == the reference's scipy Hungarian value
    (fixture made by the reference's own function, tests/golden/make_golden.py) to 1e-6 relative."""
    from vtaco_b200.common import EarthMoverDistance
    g = load('emd.npz')
    p1, p2 = g[tag + '.p1'], g[tag + '.p2']
    got, assign = EarthMoverDistance(torch.from_numpy(p1).cuda(), p2, return_assignment=True)
    want = float(g[tag + '.emd'])
    assert abs(got - want) <= 1e-6 * want, (got, want, EarthMoverDistance.last_iterations)
    a = assign.cpu().numpy()
    used = a[a >= 0]
    assert len(used) == min(len(p1), len(p2)) and len(np.unique(used)) == len(used)      # a matching
    d = np.sqrt(((p1.astype(np.float64)[a >= 0] - p2.astype(np.float64)[used]) ** 2).sum(-1))
    assert abs(d.sum() / len(p1) - got) <= 1e-12 + 1e-9 * got                            # the value is that matching's cost

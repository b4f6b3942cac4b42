"""Generator3D.generate_obj_mesh_wnf (reference generation.py:115-284, called by train.py:246): same
signature / return triple; mesh == oracle marching cubes of the decoded grid, (emd, cd) == the oracle's
metrics on the same shuffled vertices; both tactile branches."""
import numpy as np
import pytest
import torch

from oracle import convonet as oc, marching_cubes as omc

pytestmark = pytest.mark.gpu


def _net(with_unet=False):
    from vtaco_b200.encoder import encoder_dict
    from vtaco_b200.conv_onet.models import decoder_dict, ConvolutionalOccupancyNetwork
    torch.manual_seed(0)
    enc = encoder_dict['pointnet_local_pool'](dim=3, c_dim=32, padding=0.1, hidden_dim=32, plane_type='grid',
                                              grid_resolution=32)
    dec = decoder_dict['simple_local'](dim=3, c_dim=32, padding=0.1, hidden_size=32)
    with torch.no_grad():
        for m in (enc, dec):
            for b in m.blocks:
                b.fc_1.weight.normal_(0, 0.1)
    net = ConvolutionalOccupancyNetwork(dec, enc, device='cuda').eval()
    enc.division = dec.division = 'true'
    return net, enc, dec


@pytest.mark.parametrize('branch', ['none', 'tips', 't2d'])
def test_generate_obj_mesh_wnf(branch):
    from vtaco_b200.conv_onet.generation import Generator3D
    from vtaco_b200.mcubes import keys_to_level
    net, enc, dec = _net()
    rs = np.random.RandomState(11)
    nx = 64
    cloud = torch.from_numpy(rs.uniform(-0.45, 0.45, size=(1, 2500, 3)).astype(np.float32))
    points_obj = torch.from_numpy(rs.uniform(-0.4, 0.4, size=(1, 2048, 3)).astype(np.float32))
    tips = rs.uniform(-0.3, 0.3, size=(5, 3))
    feat = rs.randn(1, 5, 32).astype(np.float32)
    touch = np.array([[1, 1, 0, 1, 1]])
    pts = [tips[t] + rs.randn(128, 3) * 0.01 for t in range(5)]
    data = {'inputs': cloud, 'points.points_obj': points_obj, 'inputs.touch_success': torch.from_numpy(touch),
            'tactile.features': torch.from_numpy(feat), 'tactile.tips': tips, 'tactile.points': pts}
    gen = Generator3D(net, device='cuda', resolution0=nx // 4, with_img=branch != 'none', encode_t2d=branch == 't2d',
                      padding=0.1, input_type='pointcloud')
    np.random.seed(123)
    mesh, emd, cd = gen.generate_obj_mesh_wnf(data)
    assert isinstance(emd, float) and isinstance(cd, float)
    # the decoded grid against the oracle (features from our encoder: scatter_mean atomics are not bit-stable)
    with torch.no_grad():
        c = net.encode_inputs(cloud.cuda())
    Wd = {k: t.cpu() for k, t in dec.state_dict().items()}
    lattice = oc.dense_grid_points(nx)
    c_all = None
    if branch == 'tips':
        c_all = oc.fingertip_c_img(lattice, tips, torch.from_numpy(feat[0]), touch[0].astype(bool), 0.05)
    elif branch == 't2d':
        c_all = oc.tactile_points_c_img(lattice, pts, torch.from_numpy(feat[0]), touch[0].astype(bool), 0.015)
        assert int((c_all.abs().sum(1) > 0).sum()) > 10
    ref = oc.eval_points(lattice, {'grid': c['grid'].contiguous().cpu()}, Wd, c_all).reshape(nx, nx, nx)
    grid = gen._grid.cpu()
    assert ((grid - ref).abs() / ref.abs().clamp(min=1)).max().item() < 1e-4
    # mesh == oracle marching cubes of that grid, rescaled with the reference's 1.1/nx
    rv, rf, _ = omc.marching_cubes(grid.numpy(), None)
    rv = omc.rescale_vertices(rv, nx)
    assert np.array_equal(np.asarray(mesh.faces), rf) and np.abs(np.asarray(mesh.vertices) - rv).max() <= 1e-6
    # metrics: same shuffle (numpy global RNG), oracle Chamfer (src/common.py:69-91) and EMD (:45-51)
    np.random.seed(123)
    sv = np.asarray(mesh.vertices).copy()
    np.random.shuffle(sv)
    sv = np.ascontiguousarray(sv[:2048], dtype=np.float32)
    cd_ref = oc.chamfer_distance_naive(points_obj, torch.from_numpy(sv)[None]).item()
    emd_ref = oc.earth_mover_distance(points_obj[0].numpy(), sv)
    assert abs(cd - cd_ref) <= 1e-5 * cd_ref and abs(emd - emd_ref) <= 1e-6 * emd_ref

"""GPU: point->cell indices must be BIT-EXACT (reference src/common.py:268-348)."""
import numpy as np
import pytest
import torch

from util import load

pytestmark = pytest.mark.gpu


def test_point_to_cell_golden_true_division():
    from vtaco_b200 import common as vc
    g = load('coords.npz')
    p = torch.from_numpy(g['p']).cuda()
    for plane in ('xz', 'xy', 'yz'):
        xy = vc.normalize_coordinate(p, 0.1, plane, division='true')
        assert np.array_equal(xy.cpu().numpy(), g['norm_' + plane])
        for R in (32, 64, 128):
            idx = vc.point_to_cell(p, R, plane, 0.1, division='true')
            assert idx.dtype == torch.int64 and idx.shape == (1, 1, p.shape[1])
            assert np.array_equal(idx.cpu().numpy(), g['idx_%s_%d' % (plane, R)])
            assert np.array_equal(vc.coordinate2index(xy, R).cpu().numpy(), g['idx_%s_%d' % (plane, R)])
    pn = vc.normalize_3d_coordinate(p, 0.1, division='true')
    assert np.array_equal(pn.cpu().numpy(), g['norm_grid'])
    for R in (32, 64, 128):
        idx = vc.point_to_cell(p, R, 'grid', 0.1, division='true', index_dtype=torch.int32)
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), g['idx_grid_%d' % R])


def test_point_to_cell_matches_aten_cuda_division():
    """default mode reproduces the reference running on a GPU (ATen CUDA multiplies by 1/d)."""
    from vtaco_b200 import common as vc
    from oracle import convonet as oc
    p = torch.from_numpy(np.random.RandomState(3).uniform(-0.6, 0.6, size=(2, 500000, 3)).astype(np.float32)).cuda()
    for R in (32, 64, 128):
        ref = oc.coordinate2index(oc.normalize_3d_coordinate(p.clone(), 0.1), R, '3d')  # torch ops on CUDA
        assert torch.equal(vc.point_to_cell(p, R, 'grid', 0.1), ref)
        for plane in ('xz', 'xy', 'yz'):
            ref = oc.coordinate2index(oc.normalize_coordinate(p.clone(), 0.1, plane), R)
            assert torch.equal(vc.point_to_cell(p, R, plane, 0.1), ref)
    # and the oracle's restatement of that convention on the CPU agrees too
    ref = oc.coordinate2index(oc.normalize_3d_coordinate(p.cpu(), 0.1, cuda_division=True), 64, '3d')
    assert torch.equal(vc.point_to_cell(p, 64, 'grid', 0.1).cpu(), ref)


def test_dense_axis_bit_exact():
    from vtaco_b200.common import dense_axis
    g = load('coords.npz')
    for nx in (8, 32, 128, 256):
        assert np.array_equal(dense_axis(nx).numpy(), g['axis_%d' % nx])

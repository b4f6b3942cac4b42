"""Coordinate helpers with the reference's names and semantics
(reference src/common.py:178-197, 268-309, 333-348), CUDA-backed.

`normalize_*` / `coordinate2index` run vtaco_point_to_cell on the tensor's device.
`division` selects how `tensor / python_scalar` is evaluated: 'cuda' (ATen CUDA:
multiply by fp32(1/d), the default — what the reference does when it runs on a
GPU) or 'true' (ATen CPU: IEEE division).  See SURVEY.md §7.2-1.
"""
import torch

from . import _abi

_DIV = {'cuda': _abi.DIV_RECIPROCAL, 'reciprocal': _abi.DIV_RECIPROCAL, 'true': _abi.DIV_TRUE, 'cpu': _abi.DIV_TRUE}


def _div_mode(division):
    try:
        return _DIV[division]
    except KeyError:
        raise ValueError("division must be 'cuda' or 'true', got %r" % (division,))


def point_to_cell(p, reso, kind, padding=0.1, division='cuda', want_coord=False, index_dtype=torch.int64):
    """p (B,T,3) -> flat cell index (B,1,T) [and normalised coords (B,T,2|3)]."""
    _abi.require_cuda(p, 'p')
    _abi.forbid_autograd(p)
    if p.dim() != 3 or p.size(2) != 3:
        raise ValueError('p must have shape (B, T, 3)')
    pc = p.contiguous()
    B, T = pc.shape[0], pc.shape[1]
    idx = torch.empty((B, 1, T), dtype=index_dtype, device=p.device)
    coord = None
    if want_coord:
        coord = torch.empty((B, T, 3 if kind == 'grid' else 2), dtype=torch.float32, device=p.device)
    i32 = idx if index_dtype == torch.int32 else None
    i64 = idx if index_dtype == torch.int64 else None
    with torch.cuda.device(p.device):
        st = _abi.lib().vtaco_point_to_cell(_abi.ptr(pc), B * T, float(padding), int(reso), _abi.KIND[kind],
                                            _div_mode(division), _abi.ptr(i32), _abi.ptr(i64), _abi.ptr(coord),
                                            _abi.stream_ptr(p.device))
    _abi.check(st, 'point_to_cell')
    return (idx, coord) if want_coord else idx


def normalize_coordinate(p, padding=0.1, plane='xz', division='cuda'):
    """reference src/common.py:268-291 -> (B,T,2) in [0, 1)."""
    if plane not in ('xz', 'xy'):
        plane = 'yz'  # the reference's `else` branch
    return point_to_cell(p, 1, plane, padding, division, want_coord=True)[1]


def normalize_3d_coordinate(p, padding=0.1, division='cuda'):
    """reference src/common.py:293-309 -> (B,T,3) in [0, 1)."""
    return point_to_cell(p, 1, 'grid', padding, division, want_coord=True)[1]


def coordinate2index(x, reso, coord_type='2d'):
    """reference src/common.py:333-348 on already-normalised coordinates.
    (Plain torch integer arithmetic: this helper is not on the fused path, which
    goes point -> cell in one kernel via `point_to_cell`.)"""
    x = (x * reso).long()
    if coord_type == '2d':
        index = x[:, :, 0] + reso * x[:, :, 1]
    elif coord_type == '3d':
        index = x[:, :, 0] + reso * (x[:, :, 1] + reso * x[:, :, 2])
    else:
        raise ValueError(coord_type)
    return index[:, None, :]


def make_3d_grid(bb_min, bb_max, shape):
    """reference src/common.py:178-197 (host tensor, x slowest / z fastest)."""
    size = shape[0] * shape[1] * shape[2]
    axes = [torch.linspace(bb_min[i], bb_max[i], shape[i]) for i in range(3)]
    pxs = axes[0].view(-1, 1, 1).expand(*shape).contiguous().view(size)
    pys = axes[1].view(1, -1, 1).expand(*shape).contiguous().view(size)
    pzs = axes[2].view(1, 1, -1).expand(*shape).contiguous().view(size)
    return torch.stack([pxs, pys, pzs], dim=1)


def dense_axis(nx, padding=0.1, device=None):
    """Axis values of the extraction lattice (1+padding)*make_3d_grid((-.5,)*3,(.5,)*3,(nx,)*3)
    (reference src/conv_onet/generation.py:119,155-157), bit-identical to the
    reference's host tensor: computed with the same torch ops on the host."""
    ax = (1 + padding) * torch.linspace(-0.5, 0.5, nx)
    return ax.to(device) if device is not None else ax

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


# --------------------------------------------------------------------------- #
# Chamfer distance (reference src/common.py:54-137) — the metric Generator3D
# computes right after mesh extraction (generation.py:281); SURVEY §8f-4.
# --------------------------------------------------------------------------- #
def _chamfer(points1, points2, want_idx):
    _abi.require_cuda(points1, 'points1')
    _abi.require_cuda(points2, 'points2')
    if points1.dim() != 3 or points2.dim() != 3 or points1.size(2) != 3 or points2.size(2) != 3:
        raise ValueError('points must have shape (B, T, 3)')
    if points1.size(0) != points2.size(0):
        raise ValueError('batch sizes differ')
    _abi.forbid_autograd(points1, points2)
    B, T1, T2 = points1.size(0), points1.size(1), points2.size(1)
    dev = points1.device
    p1, p2 = points1.contiguous(), points2.contiguous()
    d12 = torch.empty((B, T1), dtype=torch.float32, device=dev)
    d21 = torch.empty((B, T2), dtype=torch.float32, device=dev)
    i12 = torch.empty((B, T1), dtype=torch.int32, device=dev) if want_idx else None
    i21 = torch.empty((B, T2), dtype=torch.int32, device=dev) if want_idx else None
    c1 = torch.empty(B, dtype=torch.float32, device=dev)
    c2 = torch.empty(B, dtype=torch.float32, device=dev)
    L = _abi.lib()
    with torch.cuda.device(dev):
        st = L.vtaco_chamfer(_abi.ptr(p1), _abi.ptr(p2), B, T1, T2, _abi.ptr(d12), _abi.ptr(i12), _abi.ptr(d21),
                             _abi.ptr(i21), _abi.ptr(c1), _abi.ptr(c2), _abi.stream_ptr(dev))
    _abi.check(st, 'chamfer')
    return c1, c2, i12, i21


def chamfer_distance_naive(points1, points2):
    """reference src/common.py:69-91: sum of the two directed mean squared nearest-neighbour
    distances; points1 is cut to points2's length when that is < 2048, sizes must then agree."""
    if points2.size()[1] < 2048:
        points1 = points1[:, :points2.size()[1], :]
    assert points1.size() == points2.size()
    c1, c2, _, _ = _chamfer(points1, points2, False)
    return c1 + c2


def chamfer_distance_kdtree(points1, points2, give_id=False):
    """reference src/common.py:94-137; the exact neighbours a kd-tree returns, by brute force on the GPU."""
    c1, c2, i12, i21 = _chamfer(points1, points2, give_id)
    if give_id:
        return c1, c2, i12.long(), i21.long()
    return c1 + c2


def chamfer_distance(points1, points2, use_kdtree=True, give_id=False):
    """reference src/common.py:54-66."""
    if use_kdtree:
        return chamfer_distance_kdtree(points1, points2, give_id=give_id)
    return chamfer_distance_naive(points1, points2)


# --------------------------------------------------------------------------- #
# Earth-Mover distance (reference src/common.py:45-51) — the second metric of
# generate_obj_mesh_wnf (generation.py:282); SURVEY §8f-4.
# --------------------------------------------------------------------------- #
def EarthMoverDistance(points1, points2, eps=1e-9, return_assignment=False):
    """reference: d = cdist(points1, points2); d[linear_sum_assignment(d)].sum() / len(d).
    points (T,3) array-likes or tensors (moved to the GPU); the minimum-cost matching is found by
    a float64 auction algorithm on the device (csrc/emd.cu), whose cost is within T*eps of the
    Hungarian optimum scipy returns — the value agrees to ~eps.  Returns a Python float
    (and, optionally, the int32 assignment of points1's rows)."""
    import ctypes as C
    dev = points1.device if torch.is_tensor(points1) and points1.is_cuda else (
        points2.device if torch.is_tensor(points2) and points2.is_cuda else torch.device('cuda', torch.cuda.current_device()))
    p1 = torch.as_tensor(points1, dtype=torch.float32).to(dev).reshape(-1, 3).contiguous()
    p2 = torch.as_tensor(points2, dtype=torch.float32).to(dev).reshape(-1, 3).contiguous()
    L = _abi.lib()
    L.vtaco_emd_workspace_bytes.restype = C.c_int64
    L.vtaco_emd_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    L.vtaco_emd.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double,
                            C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int64), C.c_void_p]
    n1, n2 = p1.shape[0], p2.shape[0]
    nbytes = L.vtaco_emd_workspace_bytes(n1, n2)
    if nbytes < 0:
        _abi.check(int(nbytes), 'emd_workspace_bytes')
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
    assign = torch.empty(n1, dtype=torch.int32, device=dev) if return_assignment else None
    out, iters = C.c_double(0.0), C.c_int64(0)
    with torch.cuda.device(dev):
        st = L.vtaco_emd(_abi.ptr(p1), n1, _abi.ptr(p2), n2, _abi.ptr(ws), ws.numel(), float(eps), C.byref(out),
                         _abi.ptr(assign), C.byref(iters), _abi.stream_ptr(dev))
    _abi.check(st, 'emd')
    EarthMoverDistance.last_iterations = iters.value
    return (out.value, assign) if return_assignment else out.value

"""GPU marching cubes — replaces `skimage.measure.marching_cubes(value_grid,
gradient_direction='ascent')` as called by the reference
(src/conv_onet/generation.py:270); conventions in oracle/mc_tables.py."""
import ctypes as C

import torch

from . import _abi

INT32_MAX, INT32_MIN = 2 ** 31 - 1, -2 ** 31


def new_minmax_key(device):
    """int32[2] accumulator for the decoder's min/max tracking."""
    return torch.tensor([INT32_MAX, INT32_MIN], dtype=torch.int32, device=device)


def keys_to_level(keys):
    """level=None of skimage: 0.5*(min+max) in fp32 (host helper; syncs)."""
    L = _abi.lib()
    k = keys.cpu().reshape(-1, 2)
    lo, hi = L.vtaco_key_to_float_host(int(k[:, 0].min())), L.vtaco_key_to_float_host(int(k[:, 1].max()))
    return float(torch.tensor(0.5, dtype=torch.float32) * (torch.tensor(lo, dtype=torch.float32) +
                                                            torch.tensor(hi, dtype=torch.float32)))


class MarchingCubes(object):
    """Re-usable extractor: keeps scratch and output buffers between calls so a steady
    stream of grids of one size runs without allocation or host synchronisation other
    than the final read of the two counters."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._scratch = None
        self._verts = None
        self._faces = None
        self._counts = torch.zeros(4, dtype=torch.int64, device=self.device)
        self._keys = torch.zeros(2, dtype=torch.int32, device=self.device)

    def _ensure(self, nbytes, vcap, fcap):
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        if self._verts is None or self._verts.size(0) < vcap:
            self._verts = torch.empty((int(vcap), 3), dtype=torch.float32, device=self.device)
        if self._faces is None or self._faces.size(0) < fcap:
            self._faces = torch.empty((int(fcap), 3), dtype=torch.int32, device=self.device)

    def __call__(self, volume, level=None, level_keys=None, voffset=0.0, vscale=1.0, sync=True,
                 x_emit=None, x_origin=0, level_ptr=None, halo=None):
        """volume: (nx,ny,nz) float32 CUDA tensor (axis0 = x).  level: float, or None ->
        0.5*(min+max) (from `level_keys` if the decoder tracked them, else computed here);
        `level_ptr`: address of a device float holding the level (multi-GPU exchange).
        Returns (vertices (V,3) float32, faces (F,3) int32) views of internal buffers
        (valid until the next call) — or, with sync=False, the un-trimmed buffers and the
        device counter tensor (int64[4]: V, F, numbered vertices, -).
        Slab mode (`x_emit` rows owned, the rest of the volume are halo rows, `x_origin` = lattice
        row of volume[0]): include/vtaco_b200.h, vtaco_marching_cubes.  `halo=(device address, rows)`:
        the halo rows are not part of `volume` but are read in place from another (ny,nz)-row buffer —
        the next rank's grid, peer-mapped (vtaco_mc_args.halo_grid)."""
        _abi.require_cuda(volume, 'volume')
        if volume.dim() != 3 or not volume.is_contiguous():
            raise ValueError('volume must be a contiguous (nx,ny,nz) tensor')
        L = _abi.lib()
        nx, ny, nz = volume.shape
        halo_ptr, halo_rows = (int(halo[0]), int(halo[1])) if halo is not None else (0, 0)
        if halo_rows:
            if level_ptr is None and level is None and level_keys is None:
                raise ValueError('a separate halo needs an explicit level (the min/max scan covers `volume` only)')
            nx += halo_rows
        n = volume.numel()
        nbytes = L.vtaco_mc_scratch_bytes(nx, ny, nz)
        vcap = self._verts.size(0) if self._verts is not None else max(1024, 12 * max(nx * ny, ny * nz, nx * nz))
        fcap = self._faces.size(0) if self._faces is not None else 2 * vcap
        self._ensure(nbytes, vcap, fcap)
        a = _abi.McArgs()
        a.grid, a.nx, a.ny, a.nz = volume.data_ptr(), nx, ny, nz
        if halo_rows:
            a.halo_grid, a.halo_rows = halo_ptr, halo_rows
        st = _abi.stream_ptr(self.device)
        with torch.cuda.device(self.device):
            if level_ptr is not None:
                a.level_ptr = int(level_ptr)
            elif level is None:
                if level_keys is None:
                    _abi.check(L.vtaco_grid_minmax(_abi.ptr(volume), n, _abi.ptr(self._keys), st), 'grid_minmax')
                    level_keys = self._keys
                a.level_keys = level_keys.data_ptr()
                a.n_level_keys = max(1, level_keys.numel() // 2)
            else:
                a.level = float(level)
            a.scratch, a.scratch_bytes = self._scratch.data_ptr(), self._scratch.numel()
            a.counts = self._counts.data_ptr()
            a.voffset, a.vscale = float(voffset), float(vscale)
            a.vertices, a.vertex_capacity = self._verts.data_ptr(), self._verts.size(0)
            a.faces, a.face_capacity = self._faces.data_ptr(), self._faces.size(0)
            a.phase = 3
            if x_emit is not None:
                a.x_emit, a.x_origin = int(x_emit), int(x_origin)
            _abi.check(L.vtaco_marching_cubes(C.byref(a), st), 'marching_cubes')
            if not sync:
                return self._verts, self._faces, self._counts
            V, F = [int(v) for v in self._counts[:2].cpu()]
            if V > self._verts.size(0) or F > self._faces.size(0):
                self._ensure(nbytes, int(V * 1.25) + 16, int(F * 1.25) + 16)
                a.vertices, a.vertex_capacity = self._verts.data_ptr(), self._verts.size(0)
                a.faces, a.face_capacity = self._faces.data_ptr(), self._faces.size(0)
                a.phase = 2
                _abi.check(L.vtaco_marching_cubes(C.byref(a), st), 'marching_cubes')
        return self._verts[:V], self._faces[:F]


_extractors = {}


def marching_cubes(volume, level=None, **kw):
    """Functional form (one cached extractor per device); returns fresh tensors."""
    dev = volume.device
    ex = _extractors.get(dev)
    if ex is None:
        ex = _extractors[dev] = MarchingCubes(dev)
    v, f = ex(volume, level, **kw)
    return v.clone(), f.clone()

"""TEST INFRASTRUCTURE ONLY — CPU (numpy) restatement of the marching-cubes call of
the reference, src/conv_onet/generation.py:268-272:

    value_grid = values.reshape(nx, nx, nx)                  # axis0 = x
    vertices, faces, normals, _ = measure.marching_cubes(value_grid, gradient_direction='ascent')
    vertices -= np.array([nx/2, nx/2, nx/2], dtype=np.float32)
    vertices *= 1.1/nx

PARITY UNPINNED (see oracle/__init__.py and oracle/mc_tables.py): scikit-image is
absent; what follows its published contract:
  * level=None  ->  level = 0.5 * (volume.min() + volume.max())   (fp32 arithmetic)
  * a corner is "above" iff value > level; spacing (1,1,1), step_size 1, no mask,
    allow_degenerate=True;
  * one vertex per grid edge whose end points differ in that test, shared by the
    (up to 4) cells around the edge, placed by the Lewiner implementation's
    inverse-distance weighting  w = 1/(FLT_EPSILON + |value - level|)  in double
    (== linear interpolation up to the epsilon), in array-index coordinates
    ordered (axis0, axis1, axis2), returned as float32;
and what is this repo's own stated convention (skimage's Lewiner tables differ on
the ambiguous cases): the triangulation table of oracle/mc_tables.py, vertex order
= lattice order of the owning point then axis, face order = lattice order of the
cell then table order.

Pinned by invariants only (tests/test_mc_cpu.py): closed 2-manifold on closed
surfaces, Euler characteristic, outward orientation, vertices on the analytic
surface.
"""
import numpy as np

from . import mc_tables as T

FLT_EPSILON = float(np.finfo(np.float32).eps)


def iso_level(volume):
    """level=None of skimage.measure.marching_cubes, in the volume's dtype (fp32)."""
    v = np.asarray(volume)
    return v.dtype.type(0.5) * (v.min() + v.max())


def marching_cubes(volume, level=None):
    """-> (vertices (V,3) float32 in index coordinates, faces (F,3) int32, case (nx-1,ny-1,nz-1) uint8)."""
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    nx, ny, nz = vol.shape
    if level is None:
        level = iso_level(vol)
    level = np.float32(level)
    above = vol > level

    # ---- vertices: one per cut grid edge, owned by the edge's lower lattice point -----
    flags = np.zeros((nx, ny, nz, 3), dtype=bool)
    flags[:-1, :, :, 0] = above[:-1] != above[1:]
    flags[:, :-1, :, 1] = above[:, :-1] != above[:, 1:]
    flags[:, :, :-1, 2] = above[:, :, :-1] != above[:, :, 1:]
    vid = (np.cumsum(flags.reshape(-1), dtype=np.int64) - 1).reshape(flags.shape)  # valid where flags
    own = np.argwhere(flags)  # lexicographic (i, j, k, axis) == id order
    i, j, k, a = own.T
    step = np.eye(3, dtype=np.int64)[a]
    v0 = vol[i, j, k].astype(np.float64) - float(level)
    v1 = vol[i + step[:, 0], j + step[:, 1], k + step[:, 2]].astype(np.float64) - float(level)
    w0 = 1.0 / (FLT_EPSILON + np.abs(v0))
    w1 = 1.0 / (FLT_EPSILON + np.abs(v1))
    t = w1 / (w0 + w1)
    verts = own[:, :3].astype(np.float64)
    verts[np.arange(len(a)), a] += t
    verts = verts.astype(np.float32)

    # ---- case index per cell ------------------------------------------------------------
    case = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.uint8)
    for c in range(8):
        ox, oy, oz = c & 1, (c >> 1) & 1, (c >> 2) & 1
        case |= (above[ox:nx - 1 + ox, oy:ny - 1 + oy, oz:nz - 1 + oz].astype(np.uint8) << c)

    # ---- faces: cell order, then table order --------------------------------------------
    ntri = T.TRI_COUNT[case].astype(np.int64)
    tbase = np.cumsum(ntri.reshape(-1)) - ntri.reshape(-1)
    total = int(ntri.sum())
    faces = np.zeros((total, 3), dtype=np.int32)
    cells = np.argwhere(ntri > 0)
    if len(cells):
        ci, cj, ck = cells.T
        ccase = case[ci, cj, ck]
        cbase = tbase.reshape(case.shape)[ci, cj, ck]
        for t_i in range(T.MAX_TRIS):
            sel = T.TRI_COUNT[ccase] > t_i
            if not sel.any():
                break
            for corner in range(3):
                e = T.TRI_TABLE[ccase[sel], 3 * t_i + corner].astype(np.int64)
                ax = T.EDGE_AXIS[e].astype(np.int64)
                off = T.EDGE_OFF[e].astype(np.int64)
                faces[cbase[sel] + t_i, corner] = vid[ci[sel] + off[:, 0], cj[sel] + off[:, 1], ck[sel] + off[:, 2], ax]
    return verts, faces, case


def rescale_vertices(vertices, nx, box=1.1):
    """generation.py:271-272 (fp32 arithmetic like the in-place numpy ops)."""
    v = vertices.astype(np.float32) - np.array([nx / 2, nx / 2, nx / 2], dtype=np.float32)
    v *= np.float32(box / nx)
    return v


# ------------------------------- mesh invariants -------------------------------------------
def mesh_stats(verts, faces):
    """edge manifoldness, Euler characteristic, signed volume."""
    f = faces.astype(np.int64)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    und = np.sort(e, 1)
    key = und[:, 0] * (len(verts) + 1) + und[:, 1]
    uniq, cnt = np.unique(key, return_counts=True)
    dkey = e[:, 0] * (len(verts) + 1) + e[:, 1]
    _, dcnt = np.unique(dkey, return_counts=True)
    p = verts.astype(np.float64)
    vol = np.einsum('ij,ij->i', p[f[:, 0]], np.cross(p[f[:, 1]], p[f[:, 2]])).sum() / 6.0
    used = np.unique(f)
    return {'V': len(verts), 'F': len(f), 'E': len(uniq), 'closed': bool((cnt == 2).all()),
            'oriented': bool((dcnt == 1).all()), 'euler': len(used) - len(uniq) + len(f),
            'signed_volume': float(vol), 'all_vertices_used': len(used) == len(verts)}

"""OFF mesh files — the wire format train.py:250-251 exports extracted meshes in
(`mesh.export('..._obj.off')` through trimesh) and src/utils/io.py:27-80 reads.
Host-side text IO; takes the (V,3) float / (F,3) int tensors Generator3D.extract_mesh returns."""
import numpy as np
import torch


def _np(a):
    return a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)


def export_off(path, vertices, faces, digits=8):
    """Write `OFF / nv nf 0 / x y z ... / 3 i j k ...` (the layout trimesh's OFF exporter emits)."""
    v = _np(vertices).astype(np.float64).reshape(-1, 3)
    f = _np(faces).astype(np.int64).reshape(-1, 3)
    if f.size and (f.min() < 0 or f.max() >= len(v)):
        raise ValueError('face index out of range')
    with open(path, 'w') as fp:
        fp.write('OFF\n%d %d 0\n' % (len(v), len(f)))
        if len(v):
            np.savetxt(fp, v, fmt='%%.%dg' % digits)
        if len(f):
            np.savetxt(fp, np.concatenate([np.full((len(f), 1), 3, dtype=np.int64), f], 1), fmt='%d')


def read_off(path):
    """reference src/utils/io.py:27-80 (including its ModelNet fix: counts on the `OFF` line).
    Returns (vertices float64 (V,3), faces int64 (F,3))."""
    with open(path, 'r') as fp:
        lines = [ln.strip() for ln in fp.readlines()]
    lines = [ln for ln in lines if ln]
    if not lines or lines[0][:3] not in ('OFF', 'off'):
        raise ValueError('invalid OFF file %s' % path)
    if len(lines[0]) > 3:
        parts, start = lines[0][3:].split(), 1
    else:
        parts, start = lines[1].split(), 2
    nv, nf = int(parts[0]), int(parts[1])
    v = np.array([[float(x) for x in ln.split()[:3]] for ln in lines[start:start + nv]], dtype=np.float64).reshape(-1, 3)
    faces = []
    for ln in lines[start + nv:start + nv + nf]:
        t = [int(x) for x in ln.split()]
        if t[0] != 3 or len(t) < 4:
            raise ValueError('only triangle faces are supported')
        faces.append(t[1:4])
    return v, np.array(faces, dtype=np.int64).reshape(-1, 3)

"""TEST INFRASTRUCTURE + table generator — marching-cubes case table.

PARITY UNPINNED: the reference calls skimage.measure.marching_cubes (Lewiner
variant; scikit-image is un-pinned in requirements.txt:6, not installed here,
not vendored, no wheel offline), see src/conv_onet/generation.py:270.  Its
33-case Lewiner tables cannot be restated from memory, so the triangulation is
DEFINED here by a rule that is consistent across neighbouring cells (hence
watertight) and identical to any marching-cubes table on the table-independent
parts (which cells are cut, which grid edges carry a vertex):

  corner c of a cell sits at offset (c&1, (c>>1)&1, (c>>2)&1) along (axis0,
  axis1, axis2); bit c of the case index is set iff value[corner] > level.
  On every cell face, viewed from outside the cell with corners in
  counter-clockwise order, each maximal cyclic run of "above" corners (that is
  not the whole face) is cut off by one directed segment from the grid edge where
  the run ends to the grid edge where it starts.  On the ambiguous face pattern
  (+-+-) this separates the two above corners.  Because the rule depends only on
  the four face-corner signs, the two cells sharing a face agree.
  Segments chain into closed loops over the cell's cut edges; every loop is
  triangulated (first fan, by apex position, that has no diagonal lying in a cell
  face — such a diagonal could be duplicated by the neighbouring cell and make the
  mesh non-manifold — else the first such general triangulation).  With this direction the right-hand normal would point to the
  above side; triangles are emitted REVERSED so that normals point to the lower
  values, i.e. out of an object whose inside has the higher logits
  (gradient_direction='ascent' in the reference's call).

Edge numbering: e = 4*a + j for the edge along axis a whose lower corner has the
other two offsets (o1, o2) in increasing axis order, j = o1 + 2*o2.

Running this file regenerates vtaco_b200/csrc/mc_tables.h.
"""
import os

import numpy as np

MAX_TRIS = None  # filled below


def corner_of(x, y, z):
    return x | (y << 1) | (z << 2)


def edge_of(c0, c1):
    d = c0 ^ c1
    a = {1: 0, 2: 1, 4: 2}[d]
    lo = min(c0, c1)
    off = [(lo >> k) & 1 for k in range(3)]
    others = [k for k in range(3) if k != a]
    return 4 * a + off[others[0]] + 2 * off[others[1]]


def edge_owner(e):
    """(axis, (ox,oy,oz)) : lattice offset of the point that owns edge e of a cell."""
    a, j = divmod(e, 4)
    others = [k for k in range(3) if k != a]
    off = [0, 0, 0]
    off[others[0]] = j & 1
    off[others[1]] = j >> 1
    return a, tuple(off)


def faces():
    """6 faces, each a list of 4 corner ids counter-clockwise seen from outside."""
    out = []
    for a in range(3):
        b, c = (a + 1) % 3, (a + 2) % 3
        for s in (0, 1):
            ring = [(0, 0), (1, 0), (1, 1), (0, 1)]
            if s == 0:
                ring = ring[::-1]
            f = []
            for (ub, uc) in ring:
                off = [0, 0, 0]
                off[a], off[b], off[c] = s, ub, uc
                f.append(corner_of(*off))
            out.append(f)
    return out


FACES = faces()


def case_loops(case):
    """closed loops of cut-edge ids for one case (direction: above side on the left
    seen from outside)."""
    nxt = {}
    for f in FACES:
        s = [(case >> c) & 1 for c in f]
        if sum(s) in (0, 4):
            continue
        for i in range(4):
            # a run of above corners starts at i if s[i] and not s[i-1]
            if s[i] and not s[i - 1]:
                j = i
                while s[(j + 1) % 4]:
                    j += 1
                e_start = edge_of(f[j % 4], f[(j + 1) % 4])   # leaving the run (ccw)
                e_end = edge_of(f[i - 1], f[i])               # entering the run
                assert e_start not in nxt
                nxt[e_start] = e_end
    loops, seen = [], set()
    for e0 in sorted(nxt):
        if e0 in seen:
            continue
        loop, e = [], e0
        while e not in seen:
            seen.add(e)
            loop.append(e)
            e = nxt[e]
        assert e == e0, 'open chain'
        loops.append(loop)
    return loops


def _face_sets():
    out = []
    for f in FACES:
        out.append({edge_of(f[i], f[(i + 1) % 4]) for i in range(4)})
    return out


FACE_EDGES = _face_sets()


def _coplanar(e1, e2):
    """both cut edges lie on one cell face: a triangle side joining them lies in that face
    plane, where the neighbouring cell may create the same side (non-manifold edge)."""
    return any(e1 in fs and e2 in fs for fs in FACE_EDGES)


def _triangulations(poly):
    """all triangulations of a convex-position polygon given as a vertex list."""
    if len(poly) < 3:
        yield []
        return
    if len(poly) == 3:
        yield [tuple(poly)]
        return
    a, b = poly[0], poly[-1]
    for m in range(1, len(poly) - 1):
        for left in _triangulations(poly[:m + 1]):
            for right in _triangulations(poly[m:]):
                yield left + [(a, poly[m], b)] + right


def triangulate_loop(loop):
    """Fan triangulations first (apex = each loop position, in order), then every other
    triangulation; the first one without a face-coplanar diagonal wins."""
    n = len(loop)
    ring = {(loop[i], loop[(i + 1) % n]) for i in range(n)} | {(loop[(i + 1) % n], loop[i]) for i in range(n)}

    def bad(tris):
        cnt = 0
        for t in tris:
            for u, v in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                if (u, v) not in ring and _coplanar(u, v):
                    cnt += 1
        return cnt

    cands = []
    for r in range(n):
        rot = loop[r:] + loop[:r]
        cands.append([(rot[0], rot[i], rot[i + 1]) for i in range(1, n - 1)])
    best = min(cands, key=bad)
    if bad(best) == 0:
        return best
    for tris in _triangulations(loop):
        if bad(tris) == 0:
            return tris
    return best


def case_triangles(case):
    tris = []
    for loop in case_loops(case):
        for (a, b, c) in triangulate_loop(loop):
            tris.append((a, c, b))  # reversed: normals toward lower values
    return tris


def build_tables():
    tri = [case_triangles(c) for c in range(256)]
    max_t = max(len(t) for t in tri)
    table = -np.ones((256, max_t * 3), dtype=np.int8)
    count = np.zeros(256, dtype=np.int8)
    edge_mask = np.zeros(256, dtype=np.int16)
    for c in range(256):
        count[c] = len(tri[c])
        for t, (a, b, d) in enumerate(tri[c]):
            table[c, 3 * t:3 * t + 3] = (a, b, d)
            edge_mask[c] |= (1 << a) | (1 << b) | (1 << d)
    return table, count, edge_mask


TRI_TABLE, TRI_COUNT, EDGE_MASK = build_tables()
MAX_TRIS = TRI_TABLE.shape[1] // 3
EDGE_AXIS = np.array([edge_owner(e)[0] for e in range(12)], dtype=np.int8)
EDGE_OFF = np.array([edge_owner(e)[1] for e in range(12)], dtype=np.int8)


def write_header(path):
    lines = ['// GENERATED by oracle/mc_tables.py — do not edit.  See that file for the rule.',
             '#pragma once', '#include <stdint.h>', 'namespace vtaco {',
             'constexpr int kMcMaxTris = %d;' % MAX_TRIS,
             '// global memory (not __constant__): the kernels stage the tables in shared memory with coalesced loads;',
             '// every lane indexes them with its own case, which a constant bank would serialise per distinct address',
             '__device__ const int8_t kMcTriCount[256] = {%s};' % ','.join(str(int(v)) for v in TRI_COUNT),
             '// rows padded to 16 bytes: a case row is one aligned 16-byte load',
             'alignas(16) __device__ const int8_t kMcTriTable[256][16] = {']
    for c in range(256):
        lines.append('  {%s,0},' % ','.join(str(int(v)) for v in TRI_TABLE[c]))
    lines.append('};')
    lines.append('// edge e: axis, owner offset (x,y,z)')
    lines.append('alignas(16) __device__ const int8_t kMcEdge[12][4] = {')
    for e in range(12):
        lines.append('  {%d,%d,%d,%d},' % (int(EDGE_AXIS[e]), *[int(v) for v in EDGE_OFF[e]]))
    lines.append('};')
    lines.append('}  // namespace vtaco')
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')


def header_path():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return os.path.join(root, 'vtaco_b200', 'csrc', 'mc_tables.h')


if __name__ == '__main__':
    write_header(header_path())
    print('max triangles per cell:', MAX_TRIS, 'header:', header_path())

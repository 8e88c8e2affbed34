"""Iso-surface extraction for `nerf_sh.gen_mesh` (nerf_sh/gen_mesh.py:84-156).  The reference hands the sigma grid
to PyMCubes (`mcubes.marching_cubes`, a third-party package this image does not carry); here the surface comes from
marching TETRAHEDRA over the same grid: every cell is cut into the six tetrahedra that share its main diagonal
(Kuhn's subdivision — neighbouring cells cut their common face along the same diagonal, so the mesh is watertight
without a 256-case table), every tetrahedron contributes 0, 1 or 2 triangles, vertices sit on grid edges by linear
interpolation exactly like marching cubes and are shared between triangles.  Output convention is PyMCubes':
vertices in grid-index coordinates, faces wound so that normals point from inside (value > iso) to outside.

Host numpy: a 300^3 grid has ~10^5..10^6 surface cells, and only those are processed.
"""
import itertools

import numpy as np


def marching_tetrahedra(vol, iso):
    """vol [nx,ny,nz] float, iso float -> (vertices [V,3] float64 in index coordinates, triangles [F,3] int64)."""
    vol = np.asarray(vol)
    if vol.ndim != 3 or min(vol.shape) < 2:
        raise ValueError("vol must be a 3-d grid with at least 2 samples per axis")
    nx, ny, nz = vol.shape
    flat = vol.reshape(-1).astype(np.float64)
    strides = (ny * nz, nz, 1)
    inside_all = vol > iso
    # cells crossed by the surface: not all 8 corners on the same side
    cnt = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int8)
    for dx, dy, dz in itertools.product((0, 1), repeat=3):
        cnt += inside_all[dx:nx - 1 + dx, dy:ny - 1 + dy, dz:nz - 1 + dz]
    ci, cj, ck = np.nonzero((cnt > 0) & (cnt < 8))
    base = ci.astype(np.int64) * strides[0] + cj.astype(np.int64) * strides[1] + ck.astype(np.int64)
    edges_a, edges_b, in_side = [], [], []       # per triangle corner: the grid edge (a inside-side flag kept apart)
    for perm in itertools.permutations(range(3)):
        offs = np.concatenate([[0], np.cumsum([strides[a] for a in perm])])
        corners = base[:, None] + offs[None, :]                               # [cells, 4] global sample ids
        ins = flat[corners] > iso
        n_in = ins.sum(1)
        for lone_inside in (True, False):                                     # 1 | 3 split: one triangle
            sel = n_in == (1 if lone_inside else 3)
            if not sel.any():
                continue
            c, m = corners[sel], ins[sel] if lone_inside else ~ins[sel]
            lone = np.argmax(m, axis=1)
            rest = np.argsort(m, axis=1, kind="stable")[:, :3]               # the three others, in corner order
            lv = np.take_along_axis(c, lone[:, None], 1)
            rv = np.take_along_axis(c, rest, 1)
            edges_a.append(np.broadcast_to(lv, rv.shape))
            edges_b.append(rv)
            in_side.append(np.full(len(c), lone_inside))
        sel = n_in == 2                                                       # 2 | 2 split: a quad, two triangles
        if sel.any():
            c, m = corners[sel], ins[sel]
            order = np.argsort(~m, axis=1, kind="stable")                     # inside pair first, then outside pair
            p = np.take_along_axis(c, order, 1)                               # columns: in0, in1, out0, out1
            quad_a = np.stack([p[:, 0], p[:, 0], p[:, 1], p[:, 1]], 1)        # around the quad: (i0,o0) (i0,o1)
            quad_b = np.stack([p[:, 2], p[:, 3], p[:, 3], p[:, 2]], 1)        #                  (i1,o1) (i1,o0)
            for tri in ((0, 1, 2), (0, 2, 3)):
                edges_a.append(quad_a[:, tri])
                edges_b.append(quad_b[:, tri])
                in_side.append(np.ones(len(c), dtype=bool))
    if not edges_a:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    ea, eb = np.concatenate(edges_a), np.concatenate(edges_b)                 # [F,3] each: edge = (ea, eb)
    a_inside = np.concatenate(in_side)                                        # True: ea is the inside end
    # one vertex per distinct grid edge, interpolated from its lower-numbered end (bit-identical for all users)
    lo, hi = np.minimum(ea, eb), np.maximum(ea, eb)
    key = lo * np.int64(flat.size) + hi
    uniq, first, inverse = np.unique(key.reshape(-1), return_index=True, return_inverse=True)
    ulo, uhi = lo.reshape(-1)[first], hi.reshape(-1)[first]
    t = (iso - flat[ulo]) / (flat[uhi] - flat[ulo])
    plo = np.stack(np.unravel_index(ulo, vol.shape), 1).astype(np.float64)
    phi = np.stack(np.unravel_index(uhi, vol.shape), 1).astype(np.float64)
    verts = plo + t[:, None] * (phi - plo)
    tris = inverse.reshape(-1, 3).astype(np.int64)
    # orientation: normal along (outside end - inside end) of the triangle's edges
    pa = np.stack(np.unravel_index(ea.reshape(-1), vol.shape), 1).reshape(-1, 3, 3).astype(np.float64)
    pb = np.stack(np.unravel_index(eb.reshape(-1), vol.shape), 1).reshape(-1, 3, 3).astype(np.float64)
    outward = (pb - pa).sum(1) * np.where(a_inside, 1.0, -1.0)[:, None]
    v0, v1, v2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    normal = np.cross(v1 - v0, v2 - v0)
    flip = (normal * outward).sum(1) < 0
    tris[flip] = tris[flip][:, ::-1]
    degenerate = (tris[:, 0] == tris[:, 1]) | (tris[:, 1] == tris[:, 2]) | (tris[:, 0] == tris[:, 2])
    return verts, tris[~degenerate]


def save_obj(vertices, triangles, path, vert_rgb=None):
    """Wavefront OBJ, optionally with per-vertex colours (gen_mesh.py:134-156: `v x y z [r g b]`, 1-based `f`)."""
    vertices, triangles = np.asarray(vertices), np.asarray(triangles)
    with open(path, "w") as f:
        if vert_rgb is None:
            f.writelines("v %.4f %.4f %.4f\n" % tuple(v) for v in vertices)
        else:
            f.writelines("v %.4f %.4f %.4f %.4f %.4f %.4f\n" % (*v, *c) for v, c in zip(vertices, np.asarray(vert_rgb)))
        f.writelines("f %d %d %d\n" % tuple(t + 1) for t in triangles)

"""Synthetic P1 triangle meshes for the BASELINE.json configs.

Stands in for ``gmshio.read_from_msh`` (reference setups/setup_cooke2.py:19): returns the two
arrays the solver consumes, vertex coordinates ``xy (Nv,2) float64`` and the cell->vertex map
``cells (Ne,3) int32``.  Geometry-node index == P1 dof index, as the reference assumes
(source/model_setup.py:70-71,85-89).
"""
import numpy as np


def rectangle(nx, ny, lx, ly, x0=0.0, y0=0.0, jitter=0.0, seed=1234, diagonal="right"):
    """``nx`` x ``ny`` quads on [x0,x0+lx] x [y0,y0+ly], each split in two triangles.

    jitter: interior vertices are moved by U(-jitter*h, jitter*h) (h = cell size per axis).
    diagonal: "right" (all /), "left" (all \\), "alternate" (checkerboard) or "random".
    """
    xs = np.linspace(x0, x0 + lx, nx + 1)
    ys = np.linspace(y0, y0 + ly, ny + 1)
    X, Y = np.meshgrid(xs, ys)                       # (ny+1, nx+1), vertex id = iy*(nx+1)+ix
    xy = np.stack([X.ravel(), Y.ravel()], axis=1)
    rng = np.random.default_rng(seed)
    if jitter > 0.0:
        hx, hy = lx / nx, ly / ny
        # one (dx, dy) pair per interior vertex, drawn in vertex order (row-major over the interior block)
        d = rng.uniform(-jitter, jitter, size=(max(ny - 1, 0) * max(nx - 1, 0), 2))
        if d.size:
            d = d.reshape(ny - 1, nx - 1, 2)
            g = xy.reshape(ny + 1, nx + 1, 2)
            g[1:-1, 1:-1, 0] += d[:, :, 0] * hx
            g[1:-1, 1:-1, 1] += d[:, :, 1] * hy
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny))
    v00 = (iy * (nx + 1) + ix).ravel()
    v10 = v00 + 1
    v01 = v00 + nx + 1
    v11 = v01 + 1
    if diagonal == "right":
        flip = np.zeros(v00.size, dtype=bool)
    elif diagonal == "left":
        flip = np.ones(v00.size, dtype=bool)
    elif diagonal == "alternate":
        flip = ((ix + iy) % 2 == 1).ravel()
    elif diagonal == "random":
        flip = rng.random(v00.size) < 0.5
    else:
        raise ValueError(f"unknown diagonal {diagonal!r}")
    # right: (v00,v10,v11),(v00,v11,v01); left: (v00,v10,v01),(v10,v11,v01)
    cells = np.empty((2 * v00.size, 3), dtype=np.int32)
    if not flip.any():                                # written column by column: no (n,3) int64 temporaries
        cols0, cols1 = (v00, v10, v11), (v00, v11, v01)
    else:
        cols0 = (v00, v10, np.where(flip, v01, v11))
        cols1 = (np.where(flip, v10, v00), v11, v01)
    for k in range(3):
        cells[0::2, k] = cols0[k]
        cells[1::2, k] = cols1[k]
    return xy, cells


def scramble(xy, cells, seed=7):
    """Random renumbering of vertices and cells plus random local rotations/reflections:
    an 'as delivered by a mesher' ordering for tests (gmsh-like, no locality)."""
    rng = np.random.default_rng(seed)
    nv, ne = xy.shape[0], cells.shape[0]
    pv = rng.permutation(nv)                 # new id of old vertex
    xy2 = np.empty_like(xy)
    xy2[pv] = xy
    c = pv[cells]
    c = c[rng.permutation(ne)]
    rot = rng.integers(0, 3, size=ne)
    c = np.stack([c[np.arange(ne), (rot + k) % 3] for k in range(3)], axis=1)
    ref = rng.random(ne) < 0.5
    c[ref] = c[ref][:, [0, 2, 1]]
    return xy2, np.ascontiguousarray(c, dtype=np.int32)

"""CPU marching cubes over the generated case table (TEST INFRASTRUCTURE ONLY): the restatement the CUDA kernels of
csrc/mcubes.cu are checked against, and the place where the table's topology (watertight, consistently oriented) is
proved on random smooth fields.  `mcubes.marching_cubes(u, threshold)` (reference renderer.py:36) is a third-party host
library that is not installed here; its contract - vertices in grid-index coordinates [V,3] float, triangles [F,3] int,
the piecewise-linear isosurface of u at `threshold` with linear interpolation along cut grid edges - is what this
follows."""
from __future__ import annotations

import numpy as np

from vdn_nerf_b200.mcubes_table import CORNERS, EDGE_AXIS, TRI_COUNT, TRI_TABLE


def marching_cubes(u: np.ndarray, threshold: float = 0.0):
    """(vertices [V,3] float64 in index coordinates, triangles [F,3] int64); inside = u > threshold."""
    nx, ny, nz = u.shape
    inside = u > threshold
    case = np.zeros((nx - 1, ny - 1, nz - 1), dtype=np.int64)
    for c in range(8):
        dx, dy, dz = CORNERS[c]
        case |= inside[dx: nx - 1 + dx, dy: ny - 1 + dy, dz: nz - 1 + dz].astype(np.int64) << c
    cells = np.argwhere(TRI_COUNT[case] > 0)
    keys, pos = [], []
    for (i, j, k) in cells:
        m = case[i, j, k]
        for t in range(TRI_COUNT[m]):
            for e in TRI_TABLE[m, 3 * t: 3 * t + 3]:
                a, axis = EDGE_AXIS[e]
                p0 = np.array([i, j, k]) + CORNERS[a]
                p1 = p0.copy()
                p1[axis] += 1
                u0, u1 = float(u[tuple(p0)]), float(u[tuple(p1)])
                tt = (threshold - u0) / (u1 - u0)
                keys.append((int(p0[0]) * ny * nz + int(p0[1]) * nz + int(p0[2])) * 3 + axis)
                pos.append(p0 + tt * (p1 - p0))
    if not keys:
        return np.zeros((0, 3)), np.zeros((0, 3), dtype=np.int64)
    keys = np.array(keys, dtype=np.int64)
    pos = np.array(pos, dtype=np.float64)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    return pos[first], inv.reshape(-1, 3)


def mesh_report(verts, tris):
    """Topology / orientation numbers of an indexed triangle mesh: (#edges with != 2 incident triangles, #directed edges
    used more than once, Euler characteristic, signed volume)."""
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    und = np.sort(e, axis=1)
    _, cnt = np.unique(und, axis=0, return_counts=True)
    _, dcnt = np.unique(e, axis=0, return_counts=True)
    p0, p1, p2 = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
    vol = np.einsum("ij,ij->i", p0, np.cross(p1, p2)).sum() / 6.0
    return int((cnt != 2).sum()), int((dcnt > 1).sum()), int(len(verts) - len(cnt) + len(tris)), float(vol)

"""Marching-cubes case table, GENERATED (not transcribed): for each of the 256 corner-sign cases the triangles as triples
of cube-edge indices, padded with -1 to 16 entries like the classic table.

Construction (Lorensen & Cline's surface, with a face-consistent resolution of the ambiguous faces so that neighbouring
cubes always agree and the mesh is watertight):
  * corners c = x + 2 y + 4 z of the unit cube; the 12 edges join corners differing in one coordinate;
  * an edge is CUT when exactly one endpoint is inside (bit set);
  * on each of the 6 faces the cut edges are joined pairwise: two cut edges -> one segment; four (inside / outside corners
    alternate around the face) -> two segments, each cutting off ONE INSIDE corner - a rule that depends only on the four
    corner signs of the face, hence identical seen from both cubes that share it;
  * every cut edge lies on two faces, so the segments close into loops; each loop is triangulated as a fan and oriented
    so that its normal points from the inside corners to the outside corners.
`mcubes` (the reference's third-party dependency, renderer.py:36) uses the same construction principle; its triangle
ORDER may differ, the surface is the same piecewise-linear isosurface up to the choice on ambiguous faces.
"""
from __future__ import annotations

import numpy as np

CORNERS = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], dtype=np.int64)
EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1]      # 12 edges, a < b
EDGE_ID = {e: i for i, e in enumerate(EDGES)}
# edge -> (corner a, axis): the edge runs from corner a along +axis
EDGE_AXIS = [(a, int(np.log2(a ^ b))) for a, b in EDGES]


def _faces():
    faces = []
    for axis in range(3):
        for side in (0, 1):
            cs = [c for c in range(8) if ((c >> axis) & 1) == side]
            # cyclic order around the face: walk corners so that consecutive ones differ in one bit
            u, v = [a for a in range(3) if a != axis]
            base = side << axis
            cyc = [base, base | (1 << u), base | (1 << u) | (1 << v), base | (1 << v)]
            assert sorted(cyc) == sorted(cs)
            faces.append(cyc)
    return faces


FACES = _faces()


def _edge(a, b):
    return EDGE_ID[(min(a, b), max(a, b))]


def _case(mask: int):
    inside = [(mask >> c) & 1 for c in range(8)]
    adj = {}                     # cut edge -> list of neighbouring cut edges (via face segments)

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)
    for cyc in FACES:
        cut = [(i, _edge(cyc[i], cyc[(i + 1) % 4])) for i in range(4) if inside[cyc[i]] != inside[cyc[(i + 1) % 4]]]
        if len(cut) == 2:
            link(cut[0][1], cut[1][1])
        elif len(cut) == 4:
            # corners alternate; cut off each INSIDE corner: join the two face edges incident to it
            for i in range(4):
                if inside[cyc[i]]:
                    link(_edge(cyc[i - 1], cyc[i]), _edge(cyc[i], cyc[(i + 1) % 4]))
    tris = []
    seen = set()
    mid = lambda e: (CORNERS[EDGES[e][0]] + CORNERS[EDGES[e][1]]) / 2.0
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxts = [n for n in adj[cur] if n != prev]
            nxt = nxts[0] if nxts else adj[cur][0]
            if len(adj[cur]) == 2 and adj[cur][0] == adj[cur][1]:
                nxt = adj[cur][0]
            if nxt == start:
                break
            if nxt in seen:          # degenerate two-edge loop
                break
            loop.append(nxt)
            seen.add(nxt)
            prev, cur = cur, nxt
        if len(loop) < 3:
            continue
        # orientation: Newell normal vs the inside -> outside direction summed over the loop's cut edges
        pts = np.array([mid(e) for e in loop])
        nrm = np.zeros(3)
        for i in range(len(loop)):
            p, q = pts[i], pts[(i + 1) % len(loop)]
            nrm += np.cross(p, q)
        out_dir = np.zeros(3)
        for e in loop:
            a, b = EDGES[e]
            d = (CORNERS[b] - CORNERS[a]).astype(float)
            out_dir += d if inside[a] else -d
        if np.dot(nrm, out_dir) < 0:
            loop = loop[::-1]
        tris.extend(_triangulate(loop))
    return tris


def _share_face(e0, e1):
    c = set(EDGES[e0]) | set(EDGES[e1])
    return any(c <= set(f) for f in FACES)


def _triangulations(idx):
    """All triangulations of the convex polygon with vertex positions idx (tuples of index triples)."""
    if len(idx) < 3:
        yield []
        return
    if len(idx) == 3:
        yield [tuple(idx)]
        return
    a, b = idx[0], idx[-1]
    for k in range(1, len(idx) - 1):
        for left in _triangulations(idx[: k + 1]):
            for right in _triangulations(idx[k:]):
                yield left + [(a, idx[k], b)] + right


def _triangulate(loop):
    """Triangles of one loop.  A diagonal that joins two cut edges of the SAME cube face would coincide with a possible
    face segment (of another loop, or of the neighbouring cube) and make the mesh non-manifold: choose a triangulation
    whose diagonals all run through the cube's interior."""
    n = len(loop)
    best = None
    for tri in _triangulations(list(range(n))):
        bad = 0
        for (i, j, k) in tri:
            for p, q in ((i, j), (j, k), (k, i)):
                if (q - p) % n not in (1, n - 1) and _share_face(loop[p], loop[q]):
                    bad += 1
        if best is None or bad < best[0]:
            best = (bad, tri)
        if bad == 0:
            break
    # keep the loop's orientation: (i, j, k) with i < j < k is counter-clockwise in loop order
    return [tuple(loop[v] for v in sorted(t)) for t in best[1]]


def build_table():
    table = -np.ones((256, 16), dtype=np.int32)
    counts = np.zeros(256, dtype=np.int32)
    for m in range(256):
        tris = _case(m)
        assert len(tris) <= 5, (m, len(tris))
        counts[m] = len(tris)
        flat = [e for t in tris for e in t]
        table[m, : len(flat)] = flat
    return table, counts


TRI_TABLE, TRI_COUNT = build_table()

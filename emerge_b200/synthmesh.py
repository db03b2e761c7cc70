"""Synthetic structured tetrahedral meshes (host side, numpy only).

The reference obtains its meshes from gmsh, which is absent here and on the GPU box.
BASELINE.json configs 4-5 name *synthetic* boxes, so this module generates them: an
nx*ny*nz grid of cells, every cell cut into 6 Kuhn tetrahedra that share the cell's main
diagonal, every tetrahedron positively oriented (the reference assumes that,
fem/mth/optimized.py:354 takes |det| but keeps signed cofactors - SURVEY App. A.4).

`mesh_tables` then derives the edge/triangle tables the element kernels consume, in the
layout of fem/mesh3d.py:252-352 (sorted vertex tuples; local edge order 1-2,1-3,1-4,2-3,
4-2,3-4; local face order 1-2-3,1-3-4,1-4-2,2-3-4).  The *numbering* of edges/triangles here
is lexicographic (np.unique), not the reference's CPython-set order; parity tests therefore
consume the reference's own tables from tests/golden, and this builder is used where no
reference numbering exists (large synthetic benchmarks).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field

import numpy as np


@dataclass
class BoxMesh:
    nodes_xyz: np.ndarray          # (nN, 3) float64, xyz interleaved (as gmsh serves it)
    tets: np.ndarray               # (nT, 4) int64, 0-based, positively oriented
    face_tris: np.ndarray          # (nF, 3) int64 tagged triangles (boundary + internal)
    face_tag: np.ndarray           # (nF,) int64 face tag per tagged triangle
    tet_vol: np.ndarray            # (nT,) int64 volume tag per tet
    face_normal: dict = field(default_factory=dict)   # tag -> outward normal
    dims: tuple = (0.0, 0.0, 0.0)
    cells: tuple = (0, 0, 0)


def _perm_sign(p):
    s = 1
    p = list(p)
    for i in range(len(p)):
        for j in range(i + 1, len(p)):
            if p[i] > p[j]:
                s = -s
    return s


def _morton3(i, j, k):
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
        v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
        return v
    return spread(i) | (spread(j) << np.uint64(1)) | (spread(k) << np.uint64(2))


def box_mesh(nx: int, ny: int, nz: int, a: float, b: float, L: float,
             jitter: float = 0.0, seed: int = 0, vol_fn=None,
             internal_faces=None, node_order: str = "lex") -> BoxMesh:
    """Structured box [−a/2,a/2]×[−b/2,b/2]×[0,L] → Kuhn tets.

    Face tags: 1:x=-a/2  2:x=+a/2  3:y=-b/2  4:y=+b/2  5:z=0  6:z=L.
    `vol_fn(cx,cy,cz)->int tags` partitions tets into volumes (default all 1).
    `internal_faces`: list of (tag, normal, fn(cx,cy,cz)->bool) selecting interior grid-plane
    triangles by centroid (PEC strips, lumped-port plates).
    """
    xs = np.linspace(-a / 2, a / 2, nx + 1)
    ys = np.linspace(-b / 2, b / 2, ny + 1)
    zs = np.linspace(0.0, L, nz + 1)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    nodes = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)

    def nid(i, j, k):
        return (i * (ny + 1) + j) * (nz + 1) + k

    if jitter > 0:
        rng = np.random.default_rng(seed)
        I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        interior = ((I > 0) & (I < nx) & (J > 0) & (J < ny) & (K > 0) & (K < nz)).ravel()
        h = np.array([a / nx, b / ny, L / nz])
        d = (rng.random(nodes.shape) - 0.5) * 2 * jitter * h
        nodes[interior] += d[interior]

    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    tets = []
    for perm in itertools.permutations(range(3)):
        off = np.zeros(3, dtype=np.int64)
        vs = [nid(I, J, K)]
        for ax in perm:
            off = off.copy()
            off[ax] = 1
            vs.append(nid(I + off[0], J + off[1], K + off[2]))
        v = np.stack(vs, axis=1)
        if _perm_sign(perm) < 0:
            v = v[:, [0, 2, 1, 3]]
        tets.append(v)
    tets = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int64)

    # enforce positive orientation (also after jitter)
    p = nodes[tets]
    det = np.einsum("ij,ij->i", np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), p[:, 3] - p[:, 0])
    neg = det < 0
    tets[neg] = tets[neg][:, [0, 2, 1, 3]]

    # tagged triangles: boundary faces (seen once) by plane + optional internal faces
    f = np.concatenate([tets[:, [0, 1, 2]], tets[:, [0, 2, 3]], tets[:, [0, 3, 1]], tets[:, [1, 2, 3]]])
    fs = np.sort(f, axis=1)
    nN = len(nodes)
    if nN < 2_000_000:               # one int64 key per face: an order of magnitude faster than a row-wise unique
        key, cnt = np.unique((fs[:, 0] * nN + fs[:, 1]) * nN + fs[:, 2], return_counts=True)
        uniq = np.stack([key // (nN * nN), (key // nN) % nN, key % nN], axis=1)
    else:
        uniq, cnt = np.unique(fs, axis=0, return_counts=True)
    bnd = uniq[cnt == 1]
    c = nodes[bnd].mean(axis=1)
    tol = 1e-9 * max(a, b, L)
    tag = np.zeros(len(bnd), dtype=np.int64)
    tag[np.abs(c[:, 0] + a / 2) < tol] = 1
    tag[np.abs(c[:, 0] - a / 2) < tol] = 2
    tag[np.abs(c[:, 1] + b / 2) < tol] = 3
    tag[np.abs(c[:, 1] - b / 2) < tol] = 4
    tag[np.abs(c[:, 2]) < tol] = 5
    tag[np.abs(c[:, 2] - L) < tol] = 6
    assert (tag > 0).all()
    normals = {1: (-1, 0, 0), 2: (1, 0, 0), 3: (0, -1, 0), 4: (0, 1, 0), 5: (0, 0, -1), 6: (0, 0, 1)}
    face_tris, face_tag = [bnd], [tag]
    if internal_faces:
        inner = uniq[cnt == 2]
        ci = nodes[inner].mean(axis=1)
        for t, nrm, fn in internal_faces:
            sel = fn(ci[:, 0], ci[:, 1], ci[:, 2])
            face_tris.append(inner[sel])
            face_tag.append(np.full(int(sel.sum()), t, dtype=np.int64))
            normals[t] = tuple(nrm)
    face_tris = np.concatenate(face_tris)
    face_tag = np.concatenate(face_tag)

    if node_order == "morton":       # renumber the nodes along a Z-order curve (a mesher's numbering has no such guarantee)
        I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
        perm = np.argsort(_morton3(I.ravel(), J.ravel(), K.ravel()), kind="stable")
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        nodes = nodes[perm]
        tets = inv[tets]
        face_tris = inv[face_tris]
    elif node_order != "lex":
        raise ValueError(node_order)
    cen = nodes[tets].mean(axis=1)
    tet_vol = np.ones(len(tets), dtype=np.int64) if vol_fn is None else \
        np.asarray(vol_fn(cen[:, 0], cen[:, 1], cen[:, 2]), dtype=np.int64)
    return BoxMesh(nodes, tets, face_tris, face_tag, tet_vol, normals, (a, b, L), (nx, ny, nz))


@dataclass
class MeshTables:
    """The arrays of fem/mesh3d.py + fem/elements/nedelec2.py the hot path consumes."""
    nodes: np.ndarray          # (3, nN) f8
    tets: np.ndarray           # (4, nT) i8
    edges: np.ndarray          # (2, nE) i8, ascending within a column
    tris: np.ndarray           # (3, nTri) i8, ascending within a column
    tet_to_edge: np.ndarray    # (6, nT)
    tet_to_tri: np.ndarray     # (4, nT)
    tri_to_edge: np.ndarray    # (3, nTri)
    tri_to_tet: np.ndarray     # (2, nTri), -1 padded
    edge_lengths: np.ndarray   # (nE,)
    tet_to_field: np.ndarray   # (20, nT)
    tri_to_field: np.ndarray   # (8, nTri)
    edge_to_field: np.ndarray  # (2, nE)

    @property
    def n_field(self) -> int:
        return 2 * self.edges.shape[1] + 2 * self.tris.shape[1]


def _lookup(sorted_keys: np.ndarray, keys: np.ndarray) -> np.ndarray:
    idx = np.searchsorted(sorted_keys, keys)
    assert np.array_equal(sorted_keys[idx], keys)
    return idx


def mesh_tables(nodes_xyz: np.ndarray, tets_n4: np.ndarray,
                edges: np.ndarray | None = None, tris: np.ndarray | None = None) -> MeshTables:
    """Vectorised equivalent of Mesh3D.update() (fem/mesh3d.py:224-355) + Nedelec2.__init__
    (fem/elements/nedelec2.py:32-62).  Numbering: the caller's `edges` (2,nE) / `tris` (3,nTri)
    when given (e.g. the reference's set-ordered tables), else lexicographic."""
    nN = nodes_xyz.shape[0]
    T = np.asarray(tets_n4, dtype=np.int64)
    nT = T.shape[0]
    # local edge order (1-2,1-3,1-4,2-3,4-2,3-4), face order (1-2-3,1-3-4,1-4-2,2-3-4): mesh3d.py:292,296
    le = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [3, 1], [2, 3]])
    lf = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 2, 3]])
    e = np.sort(T[:, le], axis=2)                      # (nT,6,2)
    ekey = e[..., 0] * nN + e[..., 1]
    if edges is None:
        uek = np.unique(ekey.ravel())
        eperm = np.arange(len(uek))
        edges = np.stack([uek // nN, uek % nN])
    else:
        edges = np.asarray(edges, dtype=np.int64)
        k = edges[0] * nN + edges[1]
        eperm = np.argsort(k)
        uek = k[eperm]
    tet_to_edge = eperm[_lookup(uek, ekey.ravel())].reshape(nT, 6).T.copy()
    f = np.sort(T[:, lf], axis=2)                      # (nT,4,3)
    fkey = (f[..., 0] * nN + f[..., 1]) * nN + f[..., 2]
    if tris is None:
        ufk = np.unique(fkey.ravel())
        fperm = np.arange(len(ufk))
        tris = np.stack([ufk // (nN * nN), (ufk // nN) % nN, ufk % nN])
    else:
        tris = np.asarray(tris, dtype=np.int64)
        k = (tris[0] * nN + tris[1]) * nN + tris[2]
        fperm = np.argsort(k)
        ufk = k[fperm]
    tet_to_tri = fperm[_lookup(ufk, fkey.ravel())].reshape(nT, 4).T.copy()
    nE, nTri = edges.shape[1], tris.shape[1]
    # tri edges (1-2, 2-3, 1-3): mesh3d.py:330-335
    tri_to_edge = np.stack([
        eperm[_lookup(uek, tris[0] * nN + tris[1])],
        eperm[_lookup(uek, tris[1] * nN + tris[2])],
        eperm[_lookup(uek, tris[0] * nN + tris[2])]])
    tri_to_tet = np.full((2, nTri), -1, dtype=np.int64)
    order = np.argsort(tet_to_tri.T.ravel(), kind="stable")
    tri_sorted = tet_to_tri.T.ravel()[order]
    tet_sorted = order // 4
    first = np.searchsorted(tri_sorted, np.arange(nTri))
    tri_to_tet[0] = tet_sorted[first]
    cnt = np.bincount(tri_sorted, minlength=nTri)
    two = cnt == 2
    tri_to_tet[1, two] = tet_sorted[first[two] + 1]
    nodes = np.ascontiguousarray(nodes_xyz, dtype=np.float64).T     # F-order view like mesh3d.py:227
    d = nodes[:, edges[0]] - nodes[:, edges[1]]
    edge_lengths = np.sqrt((d ** 2).sum(axis=0))
    tet_to_field = np.zeros((20, nT), dtype=np.int64)
    tet_to_field[:6] = tet_to_edge
    tet_to_field[6:10] = tet_to_tri + nE
    tet_to_field[10:16] = tet_to_edge + (nTri + nE)
    tet_to_field[16:20] = tet_to_tri + (nTri + 2 * nE)
    tri_to_field = np.zeros((8, nTri), dtype=np.int64)
    tri_to_field[:3] = tri_to_edge
    tri_to_field[3] = np.arange(nTri) + nE
    tri_to_field[4:7] = tri_to_edge + nE + nTri
    tri_to_field[7] = np.arange(nTri) + 2 * nE + nTri
    edge_to_field = np.stack([np.arange(nE), np.arange(nE) + nTri + nE])
    return MeshTables(nodes, T.T, edges, tris, tet_to_edge, tet_to_tri, tri_to_edge, tri_to_tet,
                      edge_lengths, tet_to_field, tri_to_field, edge_to_field)


def tri_ids_of(tables: MeshTables, face_tris: np.ndarray) -> np.ndarray:
    """Global triangle ids of tagged triangles given as vertex triples (role of Mesh3D.get_triangles,
    fem/mesh3d.py:191-207)."""
    nN = tables.nodes.shape[1]
    key = (tables.tris[0] * nN + tables.tris[1]) * nN + tables.tris[2]
    order = np.argsort(key)
    s = np.sort(np.asarray(face_tris, dtype=np.int64), axis=1)
    k = (s[:, 0] * nN + s[:, 1]) * nN + s[:, 2]
    pos = np.searchsorted(key[order], k)
    assert np.array_equal(key[order][pos], k)
    return order[pos]

"""Mesh fixtures for the matrix-free apply path (host side, numpy only).

The reference takes its meshes from `external/smesh`, which is absent from the checkout
(SURVEY.md section 0), so the structured-cube generators and the semi-structured (sshex8)
numbering are restated here from their in-tree descriptions:

* box meshes: node id = x*nz + y*nz*nx + z, VTK hex8 corner order, hex -> 6 tets
  (python/sfem/mesh/box_mesh.py:66-67, 479-686 of the reference);
* sshex8 numbering: macro corners keep their ids, then edge nodes (unique edges in (i<j)
  lexicographic order, low -> high node id direction), then face nodes, then interior nodes
  `base + e*(L-1)^3 + (z-1)(L-1)^2 + (y-1)(L-1) + (x-1)` (legacy generator, commented out in
  mesh/sfem_hex8_mesh_graph.cpp:447-735); local lattice index z*(L+1)^2 + y*(L+1) + x
  (operators/sshex8/sshex8.hpp:18-24);
* hierarchical renumbering (coarse lattices first; mesh/sfem_hex8_mesh_graph.cpp:994-1065).

Connectivity is always SoA `elements[nxe][ne]` int32 and coordinates SoA `points[3][nn]` fp32,
which is what the reference's operator-level C functions consume.
"""
from __future__ import annotations

import numpy as np

idx_t = np.int32
geom_t = np.float32

# VTK hex8 corner -> (x, y, z) in {0,1}^3
HEX8_CORNERS = np.array(
    [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64
)
# 12 edges as corner pairs, 6 faces as corner quadruples (VTK)
HEX8_EDGES = np.array(
    [[0, 1], [1, 2], [3, 2], [0, 3], [4, 5], [5, 6], [7, 6], [4, 7], [0, 4], [1, 5], [2, 6], [3, 7]], dtype=np.int64
)
HEX8_FACES = np.array(
    [[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 5, 4], [3, 2, 6, 7], [0, 3, 7, 4], [1, 2, 6, 5]], dtype=np.int64
)


def box_points(nx: int, ny: int, nz: int, w=1.0, h=1.0, t=1.0):
    """Node coordinates of an nx*ny*nz-node box, reference numbering x*nz + y*nz*nx + z."""
    gx = np.linspace(0, w, num=nx, dtype=geom_t)
    gy = np.linspace(0, h, num=ny, dtype=geom_t)
    gz = np.linspace(0, t, num=nz, dtype=geom_t)
    # id = y*(nz*nx) + x*nz + z  -> array [ny, nx, nz]
    x = np.broadcast_to(gx[None, :, None], (ny, nx, nz)).reshape(-1).copy()
    y = np.broadcast_to(gy[:, None, None], (ny, nx, nz)).reshape(-1).copy()
    z = np.broadcast_to(gz[None, None, :], (ny, nx, nz)).reshape(-1).copy()
    return np.stack([x, y, z])


def _box_hex_corners(nx: int, ny: int, nz: int):
    """[8][ne] corner node ids of every hex cell, cells ordered zi-major, xi fastest."""
    ld = (nz, nz * nx, 1)
    zi, yi, xi = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    xi = xi.reshape(-1).astype(np.int64)
    yi = yi.reshape(-1).astype(np.int64)
    zi = zi.reshape(-1).astype(np.int64)
    out = np.empty((8, xi.size), dtype=np.int64)
    for c, (cx, cy, cz) in enumerate(HEX8_CORNERS):
        out[c] = (xi + cx) * ld[0] + (yi + cy) * ld[1] + (zi + cz) * ld[2]
    return out


def box_hex8(nx: int, ny: int, nz: int, w=1.0, h=1.0, t=1.0):
    """hex8 box: returns (elements[8][ne] int32, points[3][nn] fp32)."""
    return _box_hex_corners(nx, ny, nz).astype(idx_t), box_points(nx, ny, nz, w, h, t)


# hex -> 6 tets, as local hex corners (box_mesh.py:641-682 of the reference)
_HEX_TO_TETS = np.array(
    [[0, 1, 3, 7], [0, 1, 7, 5], [0, 4, 5, 7], [1, 2, 3, 6], [1, 3, 7, 6], [1, 5, 6, 7]], dtype=np.int64
)


def box_tet4(nx: int, ny: int, nz: int, w=1.0, h=1.0, t=1.0):
    """tet4 box (6 tets per hex cell): returns (elements[4][ne] int32, points[3][nn] fp32)."""
    hexes = _box_hex_corners(nx, ny, nz)  # [8][nh]
    nh = hexes.shape[1]
    el = np.empty((4, nh * 6), dtype=idx_t)
    for k, tet in enumerate(_HEX_TO_TETS):
        for v in range(4):
            el[v, k::6] = hexes[tet[v]]
    return el, box_points(nx, ny, nz, w, h, t)


# edge (a, b) carrying local node 4 + k of a 10-node tetrahedron (reference order: tet10 shape functions,
# operators/tet10/tet10_laplacian_inline_cpu.hpp:6-28; macro_tet4 sub-tets, operators/macro_tet4/macro_tet4_laplacian.cpp:22-66)
TET10_EDGES = ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3))


def tet4_to_tet10(elements: np.ndarray, points: np.ndarray):
    """Promote a tet4 mesh to 10-node tetrahedra (P2 / macro_tet4): one new node per unique edge at its midpoint, numbered
    after the corner nodes in order of first appearance of the sorted (lo, hi) edge key.
    Returns (elements[10][ne] int32, points[3][nn2] fp32)."""
    el = np.asarray(elements, dtype=np.int64)
    nn = points.shape[1]
    ne = el.shape[1]
    keys = np.empty((6, ne), dtype=np.int64)
    for k, (a, b) in enumerate(TET10_EDGES):
        lo, hi = np.minimum(el[a], el[b]), np.maximum(el[a], el[b])
        keys[k] = lo * nn + hi
    uniq, inv = np.unique(keys.ravel(), return_inverse=True)
    out = np.empty((10, ne), dtype=idx_t)
    out[:4] = el
    out[4:] = (nn + inv.reshape(6, ne)).astype(idx_t)
    lo, hi = uniq // nn, uniq % nn
    p64 = points.astype(np.float64)
    mid = 0.5 * (p64[:, lo] + p64[:, hi])
    return out, np.concatenate([points, mid.astype(geom_t)], axis=1)


def warp_points(points: np.ndarray) -> np.ndarray:
    """Smooth warp used by the reference's distributed apply test
    (frontend/tests/mpi/sfem_ParallelLaplacianApplyTest.cpp:22-35): keeps the mesh valid while
    making every Jacobian non-trivial."""
    x, y, z = (points[d].astype(np.float64) for d in range(3))
    px = x + 0.05 * np.sin(2 * np.pi * y) * np.sin(np.pi * z)
    py = y + 0.04 * np.sin(2 * np.pi * z) * np.sin(np.pi * x)
    pz = z + 0.03 * np.sin(2 * np.pi * x) * np.sin(np.pi * y)
    return np.stack([px, py, pz]).astype(geom_t)


def shear_points(points: np.ndarray, a=0.3, b=0.2, c=0.1) -> np.ndarray:
    """Affine shear: every hex stays exactly affine (the kernels' assumption) but FFF /
    adjugate get all off-diagonal entries."""
    x, y, z = (points[d].astype(np.float64) for d in range(3))
    return np.stack([x + a * y + b * z, y + c * z, z]).astype(geom_t)


def quadratic_field(points: np.ndarray) -> np.ndarray:
    """x^2 + y^2/2 - z^2/4 + xy/8 + yz/16 (sfem_ParallelLaplacianApplyTest.cpp:18-20)."""
    x, y, z = (points[d].astype(np.float64) for d in range(3))
    return x * x + 0.5 * y * y - 0.25 * z * z + 0.125 * x * y + 0.0625 * y * z


# --------------------------------------------------------------------------------------
# semi-structured hex8
# --------------------------------------------------------------------------------------


def sshex8_nxe(L: int) -> int:
    return (L + 1) ** 3


def sshex8_lidx(L: int, x, y, z):
    return z * (L + 1) * (L + 1) + y * (L + 1) + x


def sshex8_from_hex8(L: int, m_elements: np.ndarray, m_points: np.ndarray | None = None):
    """Refine every macro hex8 into L^3 micro hexes sharing nodes across macro elements.

    Returns (elements[(L+1)^3][ne] int32, n_nodes, interior_start, points[3][n_nodes] fp32 | None).
    Micro-node coordinates are the trilinear image of the macro corners.
    """
    m_elements = np.asarray(m_elements, dtype=np.int64)
    ne = m_elements.shape[1]
    n_macro_nodes = int(m_elements.max()) + 1
    nxe = sshex8_nxe(L)
    Lm1 = L - 1
    el = np.full((nxe, ne), -1, dtype=np.int64)
    t = np.arange(1, L, dtype=np.int64)  # interior lattice coordinate 1..L-1

    # 1) corners
    for c, (cx, cy, cz) in enumerate(HEX8_CORNERS):
        el[sshex8_lidx(L, cx * L, cy * L, cz * L)] = m_elements[c]
    base = n_macro_nodes

    # 2) edges: unique (lo, hi) pairs in lexicographic order; nodes run lo -> hi
    if Lm1 > 0:
        a = m_elements[HEX8_EDGES[:, 0]]  # [12][ne]
        b = m_elements[HEX8_EDGES[:, 1]]
        lo = np.minimum(a, b)
        hi = np.maximum(a, b)
        key = lo * n_macro_nodes + hi
        uniq, inv = np.unique(key.reshape(-1), return_inverse=True)
        edge_id = inv.reshape(12, ne)
        for k in range(12):
            c0 = HEX8_CORNERS[HEX8_EDGES[k, 0]] * L
            c1 = HEX8_CORNERS[HEX8_EDGES[k, 1]] * L
            fwd = a[k] < b[k]  # local direction c0 -> c1 agrees with lo -> hi
            for j, tj in enumerate(t):  # tj-th node from c0
                p = c0 + (c1 - c0) // L * tj
                n_from_lo = np.where(fwd, j, Lm1 - 1 - j)
                el[sshex8_lidx(L, p[0], p[1], p[2])] = base + edge_id[k] * Lm1 + n_from_lo
        base += uniq.size * Lm1

        # 3) faces: unique sorted corner quadruples, numbered in order of first appearance;
        #    canonical in-face frame: origin = smallest corner id, first axis towards the
        #    smaller of its two in-face neighbours
        f = m_elements[HEX8_FACES]  # [6][4][ne]
        fs = np.sort(f, axis=1)
        # the three smallest corner ids identify a face of a conforming hex mesh
        fkey = (fs[:, 0] * n_macro_nodes + fs[:, 1]) * n_macro_nodes + fs[:, 2]
        assert n_macro_nodes**3 < 2**62
        uq, first, finv = np.unique(fkey.reshape(-1), return_index=True, return_inverse=True)
        order = np.argsort(np.argsort(first))  # rank by first appearance
        face_id = order[finv].reshape(6, ne)
        nxf = Lm1 * Lm1
        ee = np.arange(ne)
        for k in range(6):
            q = f[k]  # [4][ne] corner ids around the face (cyclic order)
            o = np.argmin(q, axis=0)  # position of the origin in the cycle
            nxt = q[(o + 1) % 4, ee]
            prv = q[(o + 3) % 4, ee]
            first_is_next = nxt < prv
            i1 = np.where(first_is_next, (o + 1) % 4, (o + 3) % 4)  # first axis neighbour
            i2 = np.where(first_is_next, (o + 3) % 4, (o + 1) % 4)  # second axis neighbour
            lc = HEX8_CORNERS[HEX8_FACES[k]] * L  # [4][3] lattice coords of the face corners
            P0 = lc[o]  # [ne][3]
            D1 = (lc[i1] - P0) // L
            D2 = (lc[i2] - P0) // L
            for j2, t2 in enumerate(t):
                for j1, t1 in enumerate(t):
                    p = P0 + D1 * t1 + D2 * t2  # [ne][3]
                    li = sshex8_lidx(L, p[:, 0], p[:, 1], p[:, 2])
                    el[li, ee] = base + face_id[k] * nxf + j2 * Lm1 + j1
        base += uq.size * nxf

    # 4) interior
    interior_start = base
    if Lm1 > 0:
        nxv = Lm1**3
        e_ids = np.arange(ne, dtype=np.int64)
        for zi in t:
            for yi in t:
                for xi in t:
                    en = (zi - 1) * Lm1 * Lm1 + (yi - 1) * Lm1 + (xi - 1)
                    el[sshex8_lidx(L, xi, yi, zi)] = base + e_ids * nxv + en
        base += ne * nxv
    n_nodes = base
    assert el.min() >= 0

    pts = None
    if m_points is not None:
        pts = np.zeros((3, n_nodes), dtype=np.float64)
        mp = np.asarray(m_points, dtype=np.float64)
        corner = mp[:, m_elements]  # [3][8][ne]
        for zi in range(L + 1):
            for yi in range(L + 1):
                for xi in range(L + 1):
                    u, v, w = xi / L, yi / L, zi / L
                    sh = np.array(
                        [
                            (1 - u) * (1 - v) * (1 - w), u * (1 - v) * (1 - w), u * v * (1 - w), (1 - u) * v * (1 - w),
                            (1 - u) * (1 - v) * w, u * (1 - v) * w, u * v * w, (1 - u) * v * w,
                        ]
                    )
                    pts[:, el[sshex8_lidx(L, xi, yi, zi)]] = np.einsum("c,dce->de", sh, corner)
        pts = pts.astype(geom_t)
    return el.astype(idx_t), int(n_nodes), int(interior_start), pts


def sshex8_box(L: int, mx: int, my: int, mz: int, w=1.0, h=1.0, t=1.0, with_points=True):
    """sshex8 mesh of level L over an mx*my*mz macro-element box.

    Returns dict(elements, n_nodes, interior_start, points, m_elements, m_points)."""
    m_el, m_pts = box_hex8(mx + 1, my + 1, mz + 1, w, h, t)
    el, nn, istart, pts = sshex8_from_hex8(L, m_el, m_pts if with_points else None)
    return dict(elements=el, n_nodes=nn, interior_start=istart, points=pts, m_elements=m_el, m_points=m_pts, level=L)


def sshex8_derefinement_levels(L: int):
    """[1, 2, 4, ..., L] for power-of-two L, else [1, L] (smesh::derefinement_levels contract as
    used by algebra/mg/sfem_GeometricMultigrid.cpp:25-36)."""
    lv = [L]
    while lv[-1] % 2 == 0 and lv[-1] > 1:
        lv.append(lv[-1] // 2)
    if lv[-1] != 1:
        lv.append(1)
    return lv[::-1]


def sshex8_hierarchical_renumbering(L: int, elements: np.ndarray, n_nodes: int, levels=None):
    """Renumber so that the nodes of every coarser lattice form a prefix of the node range
    (mesh/sfem_hex8_mesh_graph.cpp:994-1065). Returns (new_elements, node_mapping old->new,
    n_nodes_per_level)."""
    if levels is None:
        levels = sshex8_derefinement_levels(L)
    el = np.asarray(elements, dtype=np.int64)
    ne = el.shape[1]
    mapping = np.full(n_nodes, -1, dtype=np.int64)
    corners = np.unique(
        np.concatenate([el[sshex8_lidx(L, cx * L, cy * L, cz * L)] for cx, cy, cz in HEX8_CORNERS])
    )
    mapping[corners] = corners  # macro nodes keep their ids
    next_id = int(corners.max()) + 1
    counts = [next_id]
    for l in levels[1:]:
        step = L // l
        # visit order: element-major, then z, y, x over the level-l lattice
        lat = np.array(
            [sshex8_lidx(L, xi * step, yi * step, zi * step) for zi in range(l + 1) for yi in range(l + 1) for xi in range(l + 1)]
        )
        ids = el[lat].T.reshape(-1)  # element-major
        fresh = ids[mapping[ids] < 0]
        uq, first = np.unique(fresh, return_index=True)
        uq = uq[np.argsort(first)]
        mapping[uq] = next_id + np.arange(uq.size)
        next_id += uq.size
        counts.append(next_id)
    assert mapping.min() >= 0 and next_id == n_nodes
    return mapping[el].astype(idx_t), mapping, counts


def sshex8_derefine(L: int, elements: np.ndarray, to_level: int) -> np.ndarray:
    """Connectivity of the level-`to_level` lattice embedded in a level-L one."""
    step = L // to_level
    lat = [
        sshex8_lidx(L, xi * step, yi * step, zi * step)
        for zi in range(to_level + 1)
        for yi in range(to_level + 1)
        for xi in range(to_level + 1)
    ]
    return np.ascontiguousarray(elements[lat])


def sshex8_face_nodes(L: int, elements: np.ndarray, m_points: np.ndarray, axis: int, value: float, eps=1e-6) -> np.ndarray:
    """Node ids of an sshex8 box mesh on the plane {coordinate[axis] == value}, from the connectivity and the MACRO
    corner coordinates only (no per-node coordinates: 135 M nodes x 3 floats is 1.6 GB at config C3). Only for
    axis-aligned boxes whose macro elements are axis-aligned (the box generators above)."""
    N = L + 1
    rng = np.arange(N)
    out = []
    for side in (0, L):
        fixed = [0, 0, 0]
        fixed[axis] = side
        corner = elements[sshex8_lidx(L, *fixed)]  # macro node id of that corner, per element
        sel = np.abs(m_points[axis][corner].astype(np.float64) - value) < eps
        if not sel.any():
            continue
        a, b = [d for d in range(3) if d != axis]
        for i in rng:
            for j in rng:
                c = [0, 0, 0]
                c[axis], c[a], c[b] = side, i, j
                out.append(elements[sshex8_lidx(L, *c)][sel])
    return np.unique(np.concatenate(out)).astype(np.int32) if out else np.zeros(0, dtype=np.int32)


def sshex8_face_nodes_macro(L: int, elements: np.ndarray, m_elements: np.ndarray, m_points: np.ndarray, axis: int,
                            value: float, eps=1e-6) -> np.ndarray:
    """Like sshex8_face_nodes, for connectivity whose corner rows are NOT macro node ids (e.g. the owned-first local
    numbering of sfem_b200.parallel): the macro elements on the plane are found through `m_elements` / `m_points`."""
    N = L + 1
    out = []
    a, b = [d for d in range(3) if d != axis]
    for side in (0, 1):
        k = int(np.flatnonzero((HEX8_CORNERS[:, axis] == side) & (HEX8_CORNERS[:, a] == 0) & (HEX8_CORNERS[:, b] == 0))[0])
        sel = np.abs(m_points[axis][m_elements[k]].astype(np.float64) - value) < eps
        if not sel.any():
            continue
        for i in range(N):
            for j in range(N):
                c = [0, 0, 0]
                c[axis], c[a], c[b] = side * L, i, j
                out.append(elements[sshex8_lidx(L, *c)][sel])
    return np.unique(np.concatenate(out)).astype(np.int32) if out else np.zeros(0, dtype=np.int32)


def _morton3(ix, iy, iz, bits=10):
    code = np.zeros(ix.shape, dtype=np.int64)
    for b in range(bits):
        code |= ((ix >> b) & 1) << (3 * b) | ((iy >> b) & 1) << (3 * b + 1) | ((iz >> b) & 1) << (3 * b + 2)
    return code


def sfc_reorder(elements: np.ndarray, points: np.ndarray, bits=10):
    """Reorder nodes and elements along a Morton space-filling curve -- the step the reference applies to every mesh
    before use (smesh::SFC::reorder; its source is not in the checkout, so this is the textbook Z-order on quantised
    coordinates, not a restatement). Returns (elements, points, node_permutation old->new)."""
    p = np.asarray(points, dtype=np.float64)
    lo, hi = p.min(axis=1, keepdims=True), p.max(axis=1, keepdims=True)
    q = np.minimum(((p - lo) / np.maximum(hi - lo, 1e-300) * (2**bits)).astype(np.int64), 2**bits - 1)
    node_order = np.argsort(_morton3(q[0], q[1], q[2], bits), kind="stable")
    new_id = np.empty(node_order.size, dtype=np.int64)
    new_id[node_order] = np.arange(node_order.size)
    el = new_id[np.asarray(elements, dtype=np.int64)]
    cen = p[:, np.asarray(elements, dtype=np.int64)].mean(axis=1)  # [3][ne]
    qc = np.minimum(((cen - lo) / np.maximum(hi - lo, 1e-300) * (2**bits)).astype(np.int64), 2**bits - 1)
    el = el[:, np.argsort(_morton3(qc[0], qc[1], qc[2], bits), kind="stable")]
    return np.ascontiguousarray(el.astype(idx_t)), np.ascontiguousarray(np.asarray(points)[:, node_order]), new_id

"""Boundary-value problems shared by the solver parity tests: each one is described once (mesh arrays, geometry
factors from the oracle, Dirichlet node sets) and instantiated twice --
  * `ref_*`  : numpy closures over the oracle's element kernels, to be driven by the REFERENCE's own solver headers
               (oracle/ref_solvers.py), i.e. Function::apply = sum of ops + copy of the constrained dofs
               (frontend/sfem_Function.cpp:481-496);
  * `gpu_*`  : the same objects on the device through sfem_b200.host (C++ host layer over the C ABI).
Patterns follow the reference's drivers: Poisson with u = -1 / +1 on two opposite faces and zero rhs
(drivers/poisson.exe.cpp:122-129), elasticity with (-1,0,0) / (1,0,0) on the x-faces
(ssmg/tests/sfem_SSGMGTest.cpp:158-169), 2-level mgsolve (drivers/mgsolve.cpp:170-290)."""
from __future__ import annotations

import numpy as np

from sfem_b200 import mesh


class Problem:
    """block_size, n_nodes, conditions = [(nodes int32[], component, value)], fine/coarse operator descriptions."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def n_dofs(self):
        return self.n_nodes * self.block_size

    # ---- host (oracle) side -------------------------------------------------------------------------------
    def constrained_dofs(self, conditions=None, block_size=None):
        bs = self.block_size if block_size is None else block_size
        conds = self.conditions if conditions is None else conditions
        if not conds:
            return np.zeros(0, dtype=np.int64)
        return np.concatenate([np.asarray(n, dtype=np.int64) * bs + c for n, c, _ in conds])

    def apply_constraints_host(self, x, conditions=None):
        for n, c, v in self.conditions if conditions is None else conditions:
            x[np.asarray(n, dtype=np.int64) * self.block_size + c] = v
        return x

    def ref_linear_op(self, apply_fn, conditions=None):
        """Function::apply as a `fn(x, y)` closure: y += A x, then y[c] = x[c]."""
        cd = self.constrained_dofs(conditions)

        def fn(x, y):
            y += apply_fn(np.array(x, copy=True))
            y[cd] = x[cd]

        return fn


def _face_nodes(points, axis, value, eps=1e-6):
    return np.where(np.abs(points[axis] - value) < eps)[0].astype(np.int32)


def _rhs(n, seed):
    """Generic right-hand side: makes CG iteration counts a property of the operator's spectrum instead of the
    symmetry of a plate problem (where they hinge on rounding)."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def poisson_tet4(orc, n=12, generic=False):
    """generic=False: the reference's plate problem on the unit cube (known linear answer).
    generic=True: warped coordinates (sfem_ParallelLaplacianApplyTest.cpp:22-35) and a seeded random rhs."""
    el, pts = mesh.box_tet4(n, n, n)
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    conds = [(_face_nodes(pts, 2, 0.0), 0, -1.0), (_face_nodes(pts, 2, 1.0), 0, 1.0)]
    if generic:
        pts = np.ascontiguousarray(mesh.warp_points(pts), dtype=np.float32)
    fff = orc.tet4_fff_fill(el, pts)
    nn = pts.shape[1]
    p = Problem(kind="tet4_laplacian", block_size=1, n_nodes=nn, elements=el, points=pts, fff=fff, conditions=conds,
                rhs=_rhs(nn, 7) if generic else None)
    p.ref_apply = lambda x: orc.tet4_laplacian_apply_opt(el, fff, x)
    p.ref_diag = lambda: orc.tet4_laplacian_diag_opt(el, fff, nn)
    p.exact = 2.0 * pts[2].astype(np.float64) - 1.0
    return p


def poisson_sshex8(orc, L=4, mx=3, my=2, mz=2, generic=False):
    """generic=True: sheared macro mesh (every element stays affine) and a seeded random rhs."""
    m = mesh.sshex8_box(L, mx, my, mz)
    mpts = np.ascontiguousarray(m["m_points"], dtype=np.float32)
    if generic:
        mpts = np.ascontiguousarray(mesh.shear_points(mpts), dtype=np.float32)
    fff = orc.hex8_fff_fill(m["m_elements"], mpts)
    nn = m["n_nodes"]
    pts = m["points"]
    conds = [(_face_nodes(pts, 0, 0.0), 0, -1.0), (_face_nodes(pts, 0, 1.0), 0, 1.0)]
    p = Problem(kind="sshex8_laplacian", block_size=1, n_nodes=nn, level=L, elements=m["elements"], fff=fff, m=m,
                points=pts, conditions=conds, rhs=_rhs(nn, 11) if generic else None)
    p.ref_apply = lambda x: orc.sshex8_laplacian_apply(L, m["elements"], fff, x)
    # coarse level = the hex8 macro mesh on the macro node ids (a prefix of the sshex8 numbering)
    mel = m["m_elements"]
    n_macro = int(mel.max()) + 1
    p.n_coarse_nodes = n_macro
    p.coarse_elements = mel
    p.coarse_conditions = [(n[n < n_macro], c, v) for n, c, v in conds]
    p.ref_coarse_apply = lambda x: orc.hex8_laplacian_apply_opt(mel, fff, x)
    p.exact = 2.0 * pts[0].astype(np.float64) - 1.0
    return p


def elasticity_sshex8(orc, L=4, mx=3, my=2, mz=2, mu=1.0, lam=1.0, generic=False, scale=None):
    """scale: edge length of one MACRO element (default: the unit cube split into mx x my x mz)."""
    if scale is None:
        m = mesh.sshex8_box(L, mx, my, mz)
        wx = 1.0
    else:
        m = mesh.sshex8_box(L, mx, my, mz, mx * scale, my * scale, mz * scale)
        wx = mx * scale
    mpts = np.ascontiguousarray(m["m_points"], dtype=np.float32)
    if generic:
        mpts = np.ascontiguousarray(mesh.shear_points(mpts), dtype=np.float32)
    adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
    nn = m["n_nodes"]
    pts = m["points"]
    left, right = _face_nodes(pts, 0, 0.0), _face_nodes(pts, 0, wx, eps=1e-4 * wx)
    conds = [(left, 0, -1.0), (left, 1, 0.0), (left, 2, 0.0), (right, 0, 1.0), (right, 1, 0.0), (right, 2, 0.0)]
    p = Problem(kind="sshex8_linear_elasticity", block_size=3, n_nodes=nn, level=L, elements=m["elements"], adj=adj, det=det,
                mu=mu, lam=lam, m=m, points=pts, conditions=conds, rhs=_rhs(nn * 3, 13) if generic else None)
    p.ref_apply = lambda x: orc.sshex8_linear_elasticity_apply(L, m["elements"], adj, det, mu, lam, x, nn)
    mel = m["m_elements"]
    n_macro = int(mel.max()) + 1
    p.n_coarse_nodes = n_macro
    p.coarse_elements = mel
    p.coarse_conditions = [(n[n < n_macro], c, v) for n, c, v in conds]
    p.ref_coarse_apply = lambda x: orc.hex8_linear_elasticity_apply(mel, adj, det, mu, lam, x, n_macro)
    return p


def poisson_sshex8_multilevel(orc, L=8, mx=2, my=2, mz=2):
    """The N-level hierarchy of create_ssgmg (ssmg/sfem_ssgmg.cpp, algebra/mg/sfem_GeometricMultigrid.cpp:12-78):
    levels L, L/2, ..., 1 of one sheared macro mesh. Every level is its own sshex8 mesh (own compact numbering, so the
    fast kernels apply on every level); level 1 is the hex8 macro mesh. Same Dirichlet faces on every level."""
    levels = mesh.sshex8_derefinement_levels(L)[::-1]  # finest first
    p = Problem(kind="sshex8_laplacian_multilevel", block_size=1, levels=[], rhs=None)
    m0 = mesh.sshex8_box(1, mx, my, mz)
    mpts = np.ascontiguousarray(mesh.shear_points(m0["m_points"]), dtype=np.float32)
    fff = orc.hex8_fff_fill(m0["m_elements"], mpts)
    p.fff = fff
    for l in levels:
        m = mesh.sshex8_box(l, mx, my, mz)
        pts = m["points"]  # unsheared lattice coordinates: only used to find the faces
        conds = [(_face_nodes(pts, 0, 0.0), 0, -1.0), (_face_nodes(pts, 0, 1.0), 0, 1.0)]
        p.levels.append(dict(level=l, elements=m["elements"], n_nodes=m["n_nodes"], conditions=conds, m_elements=m["m_elements"]))
    p.n_nodes = p.levels[0]["n_nodes"]
    p.conditions = p.levels[0]["conditions"]
    p.rhs = _rhs(p.n_nodes, 17)

    def ref_apply_level(i):
        lv = p.levels[i]
        if lv["level"] == 1:
            return lambda x: orc.hex8_laplacian_apply_opt(lv["m_elements"], fff, x)
        return lambda x: orc.sshex8_laplacian_apply(lv["level"], lv["elements"], fff, x)

    p.ref_apply_level = ref_apply_level
    p.ref_apply = ref_apply_level(0)
    return p


def ref_level_transfer(rs, orc, p, i):
    """(restriction, prolongation) between level i (fine) and i + 1 (coarse), each followed by apply_zero_constraints
    on its target space (sfem_GeometricMultigrid.cpp:38-64); sshex8_restrict / sshex8_prolongate of the reference."""
    fl, cl = p.levels[i], p.levels[i + 1]
    count = orc.sshex8_element_node_incidence_count(fl["level"], 1, fl["elements"], fl["n_nodes"])
    fine_cd = p.constrained_dofs(fl["conditions"])
    coarse_cd = p.constrained_dofs(cl["conditions"])
    nf, nc = fl["n_nodes"], cl["n_nodes"]

    def restrict(frm, to):
        if i == 0:
            p.ref_restrict_calls += 1
        to += orc.sshex8_restrict(fl["level"], 1, fl["elements"], count, cl["level"], 1, cl["elements"], 1,
                                  np.array(frm, copy=True), nc)
        to[coarse_cd] = 0

    def prolong(frm, to):
        to[:] = orc.sshex8_prolongate(cl["level"], 1, cl["elements"], fl["level"], 1, fl["elements"], 1,
                                      np.array(frm, copy=True), nf)
        to[fine_cd] = 0

    return rs.make_op(restrict, nc, nf), rs.make_op(prolong, nf, nc)


def gpu_level_function(p, dtype, i):
    from sfem_b200 import host, lib

    lv = p.levels[i]
    f = host.Function(dtype, lv["n_nodes"], 1)
    if lv["level"] == 1:
        op = lib.laplacian(lib.HEX8, lv["m_elements"], p.fff, lv["n_nodes"], dtype)
    else:
        op = lib.laplacian(lib.SSHEX8, lv["elements"], p.fff, lv["n_nodes"], dtype, level=lv["level"])
    f.add_operator(op)
    for nodes, comp, val in lv["conditions"]:
        f.add_dirichlet(nodes, comp, val)
    return f


def gpu_level_transfer(p, i):
    from sfem_b200 import lib

    fl, cl = p.levels[i], p.levels[i + 1]
    return lib.Transfer(fl["level"], 1, fl["elements"], fl["n_nodes"], cl["level"], 1, cl["elements"], cl["n_nodes"])


def level1_lattice(mel):
    """hex8 (VTK corner order) -> rows of a level-1 sshex8 lattice."""
    lat = np.zeros((8, mel.shape[1]), dtype=np.int32)
    for k, (cx, cy, cz) in enumerate(mesh.HEX8_CORNERS):
        lat[mesh.sshex8_lidx(1, cx, cy, cz)] = mel[k]
    return lat


# ---- reference-side object graphs (oracle/ref_solvers.RefSolvers) ---------------------------------------------
def ref_transfer_ops(rs, orc, p):
    """(restriction, prolongation) as create_gmg_data wires them (algebra/mg/sfem_GeometricMultigrid.cpp:38-64):
    hierarchical restriction followed by apply_zero_constraints on the coarse space, hierarchical prolongation
    followed by apply_zero_constraints on the fine space."""
    L, el, nn, nc, bs = p.level, p.elements, p.n_nodes, p.n_coarse_nodes, p.block_size
    count = orc.sshex8_element_node_incidence_count(L, 1, el, nn)
    fine_cd = p.constrained_dofs()
    coarse_cd = p.constrained_dofs(p.coarse_conditions)

    p.ref_restrict_calls = 0

    def restrict(frm, to):
        p.ref_restrict_calls += 1
        r = orc.sshex8_hierarchical_restriction(L, el, count, bs, np.array(frm, copy=True), nn)
        to += r[: nc * bs]
        to[coarse_cd] = 0

    def prolong(frm, to):
        full = np.zeros(nn * bs, dtype=frm.dtype)
        full[: nc * bs] = frm
        to[:] = orc.sshex8_hierarchical_prolongation(L, el, bs, full, nn)
        to[fine_cd] = 0

    return rs.make_op(restrict, nc * bs, nn * bs), rs.make_op(prolong, nn * bs, nc * bs)


# ---- device-side object graphs (sfem_b200.host) ---------------------------------------------------------------
def gpu_function(p, dtype, coarse=False):
    from sfem_b200 import host, lib

    if not coarse:
        f = host.Function(dtype, p.n_nodes, p.block_size)
        if p.kind == "tet4_laplacian":
            op = lib.laplacian(lib.TET4, p.elements, p.fff, p.n_nodes, dtype)
        elif p.kind == "sshex8_laplacian":
            op = lib.laplacian(lib.SSHEX8, p.elements, p.fff, p.n_nodes, dtype, level=p.level)
        elif p.kind == "sshex8_linear_elasticity":
            op = lib.linear_elasticity(lib.SSHEX8, p.elements, p.adj, p.det, p.mu, p.lam, p.n_nodes, dtype, level=p.level)
        else:
            raise ValueError(p.kind)
        conds = p.conditions
    else:
        f = host.Function(dtype, p.n_coarse_nodes, p.block_size)
        if p.kind == "sshex8_laplacian":
            op = lib.laplacian(lib.HEX8, p.coarse_elements, p.fff, p.n_coarse_nodes, dtype)
        else:
            op = lib.linear_elasticity(lib.HEX8, p.coarse_elements, p.adj, p.det, p.mu, p.lam, p.n_coarse_nodes, dtype)
        conds = p.coarse_conditions
    f.add_operator(op)
    for nodes, comp, val in conds:
        f.add_dirichlet(nodes, comp, val)
    return f


def gpu_transfer(p):
    from sfem_b200 import lib

    return lib.Transfer(p.level, 1, p.elements, p.n_nodes, 1, 1, level1_lattice(p.coarse_elements), p.n_coarse_nodes)

"""GPU parity of the reference's non-default formulations on the path and of the stored-element-matrix
("EM") semi-structured operators (section 8f rank 1), through the C ABI against the oracle on the same inputs:

  tet4_laplacian_apply            FFF recomputed from the points           operators/tet4/tet4_laplacian.cpp:70-135
  hex8_linear_elasticity_apply    isoparametric (SFEM_HEX8_ASSUME_AFFINE=0) operators/hex8/hex8_linear_elasticity.cpp:86-324
  sshex8_laplacian_apply          non-affine macro elements, q6            operators/sshex8/sshex8_laplacian.cpp:39-182
  affine_sshex8_laplacian_apply   element-matrix formulation               operators/sshex8/sshex8_laplacian.cpp:184-331
  ss:em:Laplacian / ss:em:LinearElasticity   sshex8_*_element_matrix + sshex8_stencil_element_matrix_apply(3)
Bars: 1e-12 (fp64) / 1e-5 (fp32)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-12, "f32": 1e-5}


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _dt(real):
    torch = _torch()
    return torch.float64 if real == "f64" else torch.float32


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _orc(oracles, real):
    o = oracles.get(("reference", real))
    if o is None:
        pytest.skip("needs the reference build (oracle/_ref)")
    return o


def _apply(op, u, dt, n_out=None):
    torch = _torch()
    x = torch.from_numpy(np.asarray(u)).to(dt).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    return x, y


def _sshex8_points(mesh, m, mpts, L):
    nn = m["n_nodes"]
    pts = np.zeros((3, nn), dtype=np.float32)
    corners = [mesh.sshex8_lidx(L, cx * L, cy * L, cz * L) for cx, cy, cz in mesh.HEX8_CORNERS]
    for k, c in enumerate(corners):
        pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
    return pts


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_tet4_laplacian_from_points(oracles, real):
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    el, pts = mesh.box_tet4(9, 7, 8)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    op = lib.laplacian_from_points(lib.TET4, el, pts, nn, _dt(real))
    for seed in (0, 1):
        u = np.random.default_rng(seed).uniform(-1, 1, nn)
        x, y = _apply(op, u, _dt(real))
        assert _rel(y.cpu().numpy(), orc.tet4_laplacian_apply_points(el, pts, u, nn)) < TOL[real]
    # constants are in the kernel; empty mesh; overwrite form
    x, y = _apply(op, np.ones(nn), _dt(real))
    assert float(y.abs().max()) < (1e-12 if real == "f64" else 1e-4)
    op.apply_overwrite(x, y)
    assert float(y.abs().max()) < (1e-12 if real == "f64" else 1e-4)
    empty = lib.laplacian_from_points(lib.TET4, el[:, :0], pts, nn, _dt(real))
    y.fill_(3.0)
    empty.apply_overwrite(x, y)
    assert float(y.abs().max()) == 0.0


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("vs", [1, 3])
def test_hex8_vector_laplacian(oracles, real, vs):
    """b200fem_vector_laplacian_create against affine_hex8_vector_laplacian_apply_fff (AoS fields of vs components)"""
    import torch

    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    if orc.kind != "reference":
        pytest.skip("hex8_vector_laplacian.cpp is compiled into the reference build of the oracle only")
    el, pts = mesh.box_hex8(7, 5, 6)
    pts = mesh.shear_points(pts)
    nn = pts.shape[1]
    fff = orc.hex8_fff_fill(el, pts)
    op = lib.vector_laplacian(lib.HEX8, el, fff, nn, vs, _dt(real))
    u = np.random.default_rng(vs).uniform(-1, 1, nn * vs)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    ref = orc.hex8_vector_laplacian_apply(el, fff, vs, u, nn)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    d = torch.zeros_like(x)
    op.diag(d)
    # the reference's diagonal recomputes the FFF from the points in its own precision: 1e-6 (as in test_hex8_diagonals)
    assert _rel(d.cpu().numpy().reshape(nn, vs)[:, 0], orc.hex8_laplacian_diag(el, pts, nn)) < 1e-5
    assert float((d.reshape(nn, vs) - d.reshape(nn, vs)[:, :1]).abs().max()) == 0.0  # every component gets the scalar diagonal


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_mass_and_lumped_mass(oracles, real):
    """b200fem_mass_create: apply = hex8_apply_mass, diag = hex8 / tet4 lumped mass, against the reference's own routines.
    The tet4 bar is 1e-10: the reference expands the determinant into 24 monomials of the coordinates (cancellation)."""
    import torch

    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    if orc.kind != "reference":
        pytest.skip("the mass routines exist in the reference build of the oracle only")
    el, pts = mesh.box_hex8(6, 5, 7)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    op = lib.mass(lib.HEX8, el, pts, nn, _dt(real))
    u = np.random.default_rng(2).uniform(-1, 1, nn)
    x, y = _apply(op, u, _dt(real))
    assert _rel(y.cpu().numpy(), orc.hex8_apply_mass(el, pts, u, nn)) < TOL[real]
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), orc.hex8_apply_mass(el, pts, u, nn)) < TOL[real]
    d = torch.zeros(nn, dtype=_dt(real), device="cuda")
    op.diag(d)
    lumped = orc.hex8_lumped_mass(el, pts, nn)
    assert _rel(d.cpu().numpy(), lumped) < TOL[real]
    # the mass of the constant function is the lumped mass, and its total the volume of the (warped) box
    x1, y1 = _apply(op, np.ones(nn), _dt(real))
    assert _rel(y1.cpu().numpy(), lumped) < TOL[real]
    tel, tpts = mesh.box_tet4(5, 4, 6)
    tpts = mesh.warp_points(tpts)
    tn = tpts.shape[1]
    top = lib.mass(lib.TET4, tel, tpts, tn, _dt(real))
    td = torch.zeros(tn, dtype=_dt(real), device="cuda")
    top.diag(td)
    assert _rel(td.cpu().numpy(), orc.tet4_lumped_mass(tel, tpts, tn)) < (1e-10 if real == "f64" else 1e-4)
    with pytest.raises(lib.B200FemError, match="hex8 only"):
        top.apply(td, torch.zeros_like(td))


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_hex8_linear_elasticity_isoparametric(oracles, real):
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    el, pts = mesh.box_hex8(6, 5, 7)
    pts = mesh.warp_points(pts)  # genuinely trilinear elements: the affine operator does not apply here
    nn = pts.shape[1]
    mu, lam = 1.3, 0.7
    op = lib.linear_elasticity_from_points(lib.HEX8, el, pts, mu, lam, nn, _dt(real))
    assert op.block_size == 3
    u = np.random.default_rng(5).uniform(-1, 1, nn * 3)
    x, y = _apply(op, u, _dt(real))
    ref = orc.hex8_linear_elasticity_apply_isoparametric(el, pts, mu, lam, u, nn)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    # on this warped mesh the affine operator is a different (approximate) one -- the isoparametric path is not an alias
    adj, det = orc.hex8_adjugate_and_det_fill(el, pts)
    assert _rel(orc.hex8_linear_elasticity_apply(el, adj, det, mu, lam, u, nn), ref) > 1e-4
    # rigid translations are in the kernel
    x, y = _apply(op, np.tile(np.array([0.3, -1.2, 0.5]), nn), _dt(real))
    assert float(y.abs().max()) < (1e-12 if real == "f64" else 1e-4)


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 4, 8])
def test_sshex8_laplacian_nonaffine(oracles, real, level):
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    m = mesh.sshex8_box(level, 3, 2, 2)
    mpts = mesh.warp_points(m["m_points"])  # trilinear macro elements
    nn = m["n_nodes"]
    pts = _sshex8_points(mesh, m, mpts, level)
    op = lib.laplacian_from_points(lib.SSHEX8, m["elements"], pts, nn, _dt(real), level=level)
    u = np.random.default_rng(level).uniform(-1, 1, nn)
    x, y = _apply(op, u, _dt(real))
    assert _rel(y.cpu().numpy(), orc.sshex8_laplacian_apply_points(level, m["elements"], pts, u)) < TOL[real]
    op.apply(x, y)  # accumulates
    assert _rel(y.cpu().numpy(), 2 * orc.sshex8_laplacian_apply_points(level, m["elements"], pts, u)) < TOL[real]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 4, 8])
def test_sshex8_em_laplacian(oracles, real, level):
    """ss:em:Laplacian: builder == sshex8_laplacian_element_matrix, apply == sshex8_stencil_element_matrix_apply, and the
    same operator == affine_sshex8_laplacian_apply (the element-matrix formulation of the affine Laplacian)."""
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    m = mesh.sshex8_box(level, 3, 2, 2, 3.0, 2.0, 2.0)
    mpts = mesh.shear_points(m["m_points"], 0.25, 0.5, 0.125)  # dyadic: the fp32 FFF we hand over are exact
    nn = m["n_nodes"]
    fff = orc.hex8_fff_fill(m["m_elements"], mpts)
    M = lib.sshex8_laplacian_element_matrix(level, fff, _dt(real))
    M_ref = orc.sshex8_laplacian_element_matrix(level, m["m_elements"], mpts)
    assert _rel(M, M_ref) < TOL[real]
    op = lib.sshex8_em(level, m["elements"], M_ref, 1, nn, _dt(real))  # the reference's matrices in, like the reference's Op
    u = np.random.default_rng(7).uniform(-1, 1, nn)
    x, y = _apply(op, u, _dt(real))
    assert _rel(y.cpu().numpy(), orc.sshex8_em_apply(level, m["elements"], M_ref, u)) < TOL[real]
    pts = _sshex8_points(mesh, m, mpts, level)
    assert _rel(y.cpu().numpy(), orc.sshex8_laplacian_apply_points(level, m["elements"], pts, u, affine=True)) < TOL[real]
    # and the default (stencil / FFF) operator computes the same thing
    default = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, _dt(real), level=level)
    _, yd = _apply(default, u, _dt(real))
    assert _rel(y.cpu().numpy(), yd.cpu().numpy()) < TOL[real]
    # diagonal = what unit applies see
    torch = _torch()
    d = torch.zeros(nn, dtype=_dt(real), device="cuda")
    op.diag(d)
    dd = torch.zeros_like(d)
    default.diag(dd)
    assert _rel(d.cpu().numpy(), dd.cpu().numpy()) < TOL[real]
    # an arbitrary (non-FFF) symmetric matrix is applied as given: the operator is not tied to the Laplacian
    rng = np.random.default_rng(8)
    R = rng.uniform(-1, 1, (m["elements"].shape[1], 8, 8))
    R = (R + R.transpose(0, 2, 1)).reshape(-1, 64)
    op2 = lib.sshex8_em(level, m["elements"], R, 1, nn, _dt(real))
    _, y2 = _apply(op2, u, _dt(real))
    assert _rel(y2.cpu().numpy(), orc.sshex8_em_apply(level, m["elements"], R, u)) < (1e-11 if real == "f64" else 1e-4)


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 4, 8])
def test_sshex8_em_linear_elasticity(oracles, real, level):
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    m = mesh.sshex8_box(level, 2, 3, 2, 2.0, 3.0, 2.0)
    mpts = mesh.shear_points(m["m_points"], 0.25, 0.5, 0.125)
    nn = m["n_nodes"]
    mu, lam = 1.3, 0.7
    adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
    K = lib.sshex8_linear_elasticity_element_matrix(level, adj, det, mu, lam, _dt(real))
    K_ref = orc.sshex8_linear_elasticity_element_matrix(level, m["m_elements"], mpts, mu, lam)
    assert _rel(K, K_ref) < TOL[real]
    op = lib.sshex8_em(level, m["elements"], K_ref, 3, nn, _dt(real))
    assert op.block_size == 3
    u = np.random.default_rng(9).uniform(-1, 1, nn * 3)
    x, y = _apply(op, u, _dt(real))
    assert _rel(y.cpu().numpy(), orc.sshex8_em_apply(level, m["elements"], K_ref, u, 3)) < TOL[real]
    # == the matrix-free macro-adjugate operator
    assert _rel(y.cpu().numpy(), orc.sshex8_linear_elasticity_apply(level, m["elements"], adj, det, mu, lam, u, nn)) < (
        1e-11 if real == "f64" else 1e-4)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_tet4_linear_elasticity(oracles, real):
    """tet4 linear elasticity: from the points like the reference's tet4_linear_elasticity_apply /
    _diag (operators/tet4/tet4_linear_elasticity.cpp:86-230), and from a cached fp32 adjugate (b200fem_adjugate_and_det_fill +
    b200fem_linear_elasticity_create with TET4; exact on dyadic geometry)."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    mu, lam = 1.3, 0.7
    el, pts = mesh.box_tet4(7, 6, 8)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    op = lib.linear_elasticity_from_points(lib.TET4, el, pts, mu, lam, nn, _dt(real))
    u = np.random.default_rng(11).uniform(-1, 1, nn * 3)
    x, y = _apply(op, u, _dt(real))
    assert _rel(y.cpu().numpy(), orc.tet4_linear_elasticity_apply_points(el, pts, mu, lam, u, nn)) < TOL[real]
    d = torch.zeros(nn * 3, dtype=_dt(real), device="cuda")
    op.diag(d)
    assert _rel(d.cpu().numpy(), orc.tet4_linear_elasticity_diag_points(el, pts, mu, lam, nn)) < TOL[real]
    x, y = _apply(op, np.tile(np.array([0.3, -1.2, 0.5]), nn), _dt(real))  # rigid translation
    assert float(y.abs().max()) < (1e-12 if real == "f64" else 1e-4)
    # cached adjugate, dyadic geometry: identical to the from-points operator
    el, pts = mesh.box_tet4(5, 6, 4, 4.0, 5.0, 3.0)
    pts = mesh.shear_points(pts, 0.25, 0.5, 0.125)
    nn = pts.shape[1]
    adj, det = lib.adjugate_and_det_fill(lib.TET4, el, pts)
    opc = lib.linear_elasticity(lib.TET4, el, adj, det, mu, lam, nn, _dt(real))
    u = np.random.default_rng(12).uniform(-1, 1, nn * 3)
    x, y = _apply(opc, u, _dt(real))
    assert _rel(y.cpu().numpy(), orc.tet4_linear_elasticity_apply_points(el, pts, mu, lam, u, nn)) < TOL[real]
    d = torch.zeros(nn * 3, dtype=_dt(real), device="cuda")
    opc.diag(d)
    assert _rel(d.cpu().numpy(), orc.tet4_linear_elasticity_diag_points(el, pts, mu, lam, nn)) < TOL[real]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", ["tet10", "macro_tet4"])
@pytest.mark.parametrize("scatter", ["patch", "atomic"])
def test_p2_tet_laplacian(oracles, real, family, scatter):
    """10-node tetrahedra: tet10_laplacian_apply_opt / _diag_opt (operators/tet10/tet10_laplacian.cpp:232-302)
    and macro_tet4_laplacian_apply_opt / _diag_opt (operators/macro_tet4/macro_tet4_laplacian.cpp:401-490), the reference's own
    routines compiled into the oracle, FFF of the 4 corners from b200fem_fff_fill."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    el, pts = mesh.box_tet4(9, 7, 8)
    pts = mesh.warp_points(pts)
    el10, pts10 = mesh.tet4_to_tet10(el, pts)
    nn = pts10.shape[1]
    et = lib.TET10 if family == "tet10" else lib.MACRO_TET4
    fff = lib.fff_fill(et, el10, pts10)
    assert _rel(fff, orc.tet4_fff_fill(el, pts)) < 1e-5
    fff = orc.tet4_fff_fill(el, pts)
    op = lib.laplacian(et, el10, fff, nn, _dt(real))
    assert op.get_option("scatter") == lib.SCATTER_ATOMIC  # the measured default of the 10-node Laplacians
    op.set_option("scatter", lib.SCATTER_PATCH if scatter == "patch" else lib.SCATTER_ATOMIC)
    assert op.get_option("patch_avg_unique_nodes") == 0  # the patch plan is built on first use
    for seed in (0, 1):
        u = np.random.default_rng(seed).uniform(-1, 1, nn)
        ref = orc.p2tet_laplacian_apply_opt(family, el10, fff, u)
        x = torch.from_numpy(u).to(_dt(real)).cuda()
        y = torch.zeros_like(x)
        op.apply(x, y)
        assert _rel(y.cpu().numpy(), ref) < TOL[real]
    op.apply(x, y)  # accumulate semantics
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    assert (op.get_option("patch_avg_unique_nodes") > 0) == (scatter == "patch")
    y.fill_(7.0)
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    ones = torch.ones(nn, dtype=_dt(real), device="cuda")
    op.apply_overwrite(ones, y)
    assert np.abs(y.cpu().numpy()).max() < (1e-12 if real == "f64" else 1e-4)
    d = torch.zeros(nn, dtype=_dt(real), device="cuda")
    op.diag(d)
    ref_d = orc.p2tet_laplacian_diag_opt(family, el10, fff, nn)
    assert _rel(d.cpu().numpy(), ref_d) < TOL[real]
    if family == "tet10":
        # the reference's tet10 diagonal is 1/6 of the diagonal of its own apply (tests/test_p2_tet_core.py pins that);
        # reference_diag = 0 gives the consistent one
        assert op.get_option("reference_diag") == 1
        op.set_option("reference_diag", 0)
        d.zero_()
        op.diag(d)
        assert _rel(d.cpu().numpy(), 6 * ref_d) < TOL[real]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", ["tet10", "macro_tet4"])
@pytest.mark.parametrize("scatter", ["patch", "atomic"])
def test_p2_tet_linear_elasticity(oracles, real, family, scatter):
    """tet10_linear_elasticity_apply_opt / _diag_opt (operators/tet10/tet10_linear_elasticity.cpp:209-355) and
    macro_tet4_linear_elasticity_apply_opt / _diag_opt (operators/macro_tet4/macro_tet4_linear_elasticity.cpp:206-277, 397-437)
    with the adjugate / determinant of the 4 corners from b200fem_adjugate_and_det_fill. tet10 fp64 bar 1e-11: the reference
    writes its face-centroid quadrature points as 0.333333333333 and sits 1.5e-12 from the exact integral this kernel computes."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    el, pts = mesh.box_tet4(7, 6, 8)
    pts = mesh.warp_points(pts)
    el10, pts10 = mesh.tet4_to_tet10(el, pts)
    nn = pts10.shape[1]
    et = lib.TET10 if family == "tet10" else lib.MACRO_TET4
    adj, det = lib.adjugate_and_det_fill(et, el10, pts10)
    mu, lam = 1.3, 0.7
    tol = (1e-11 if family == "tet10" else 1e-12) if real == "f64" else 1e-5
    op = lib.linear_elasticity(et, el10, adj, det, mu, lam, nn, _dt(real))
    op.set_option("scatter", lib.SCATTER_PATCH if scatter == "patch" else lib.SCATTER_ATOMIC)
    u = np.random.default_rng(11).uniform(-1, 1, nn * 3)
    ref = orc.p2tet_linear_elasticity_apply_opt(family, el10, adj, det, mu, lam, u)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < tol
    y.fill_(3.0)
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < tol
    d = torch.zeros(nn * 3, dtype=_dt(real), device="cuda")
    op.diag(d)
    assert _rel(d.cpu().numpy(), orc.p2tet_linear_elasticity_diag_opt(family, el10, adj, det, mu, lam, nn)) < tol


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", ["tet4", "hex8", "tet10", "macro_tet4"])
def test_packed_laplacian(oracles, real, family):
    """PackedLaplacian: b200fem_packed_laplacian_create on the arrays of a PackedMesh<uint16_t>
    (b200fem_packed_mesh_create), staged and direct kernels, against the reference's own *_laplacian_apply_opt on the packed
    connectivity expanded to global ids -- the sum PackedLaplacianApply::apply (frontend/ops/sfem_PackedLaplacian.cpp:111-176)
    forms pack by pack with the same element operators."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    if family == "hex8":
        el, pts = mesh.box_hex8(11, 9, 10)
        pts = mesh.shear_points(pts)
        et, epp = lib.HEX8, 128
    else:
        el, pts = mesh.box_tet4(9, 7, 8)
        pts = mesh.warp_points(pts)
        et, epp = lib.TET4, 256
        if family != "tet4":
            el, pts = mesh.tet4_to_tet10(el, pts)
            et = lib.TET10 if family == "tet10" else lib.MACRO_TET4
    nn = pts.shape[1]
    pm = lib.PackedMesh(el, nn, epp)
    assert pm.n_packs > 3
    g = pm.global_elements()
    ppts = np.ascontiguousarray(pts[:, pm.new_to_old])  # points in the packed numbering
    if family == "hex8":
        fff = orc.hex8_fff_fill(g, ppts)
        ref_apply = lambda u: orc.hex8_laplacian_apply_opt(g, fff, u)
    elif family == "tet4":
        fff = orc.tet4_fff_fill(g, ppts)
        ref_apply = lambda u: orc.tet4_laplacian_apply_opt(g, fff, u)
    else:
        fff = orc.tet4_fff_fill(g[:4], ppts)
        ref_apply = lambda u: orc.p2tet_laplacian_apply_opt(family, g, fff, u)
    op = lib.packed_laplacian(et, pm, fff, _dt(real))
    assert op.get_option("packed_kernel") == 1 and op.get_option("packed_staged") == 0  # the direct kernel is the measured default
    op.set_option("packed_kernel", 0)
    assert op.get_option("packed_staged") == 1  # and these packs fit shared memory
    u = np.random.default_rng(21).uniform(-1, 1, nn)
    ref = ref_apply(u)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    for direct in (0, 1):
        op.set_option("packed_kernel", direct)
        y = torch.zeros_like(x)
        op.apply(x, y)
        assert _rel(y.cpu().numpy(), ref) < TOL[real], direct
        op.apply(x, y)  # accumulate semantics (the exclusively owned nodes take a plain read-modify-write)
        assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real], direct
        y.fill_(5.0)
        op.apply_overwrite(x, y)
        assert _rel(y.cpu().numpy(), ref) < TOL[real], direct
    # the reference's atomics-free two-pass form (SFEM_PACKED_TWO_PASS, PackedLaplacianApply::apply_two_pass :178-262)
    op2 = lib.packed_laplacian(et, pm, fff, _dt(real), two_pass=True)
    assert op2.get_option("packed_two_pass") == 1 and op2.launches() == 2
    y = torch.zeros_like(x)
    op2.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    op2.apply(x, y)
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    y.fill_(5.0)
    op2.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    with pytest.raises(RuntimeError):  # selecting it without the reduce lists is refused
        op.set_option("packed_two_pass", 1)
    # the same operator as the unpacked one on the input numbering
    ones = torch.ones(nn, dtype=_dt(real), device="cuda")
    op.apply_overwrite(ones, y)
    assert np.abs(y.cpu().numpy()).max() < (1e-12 if real == "f64" else 1e-4)
    # one pack holding every element of a 30^3-node box: more nodes than shared memory takes -> the direct kernel is chosen
    if family in ("tet4", "hex8"):
        elb, ptb = (mesh.box_hex8 if family == "hex8" else mesh.box_tet4)(30, 30, 30)
        nb = ptb.shape[1]
        big = lib.PackedMesh(elb, nb, elb.shape[1])
        assert big.n_packs == 1 and 2 * big.max_nodes_per_pack * 4 > 200 * 1024
        gb = big.global_elements()
        bp = np.ascontiguousarray(ptb[:, big.new_to_old])
        fb = orc.hex8_fff_fill(gb, bp) if family == "hex8" else orc.tet4_fff_fill(gb, bp)
        opb = lib.packed_laplacian(et, big, fb, _dt(real))
        opb.set_option("packed_kernel", 0)
        assert opb.get_option("packed_staged") == 0
        ub = np.random.default_rng(22).uniform(-1, 1, nb)
        refb = orc.hex8_laplacian_apply_opt(gb, fb, ub) if family == "hex8" else orc.tet4_laplacian_apply_opt(gb, fb, ub)
        xb = torch.from_numpy(ub).to(_dt(real)).cuda()
        yb = torch.zeros_like(xb)
        opb.apply(xb, yb)
        assert _rel(yb.cpu().numpy(), refb) < TOL[real]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level,vs", [(2, 3), (4, 2), (8, 3)])
def test_sshex8_vector_laplacian(oracles, real, level, vs):
    """b200fem_sshex8_vector_laplacian_create against affine_sshex8_vector_laplacian_apply_fff
    (operators/sshex8/sshex8_vector_laplacian.cpp:16-97): AoS fields of vs components on the semi-structured lattice."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    m = mesh.sshex8_box(level, 3, 2, 2)
    mpts = mesh.shear_points(m["m_points"])
    nn = m["n_nodes"]
    fff = orc.hex8_fff_fill(m["m_elements"], mpts)
    op = lib.sshex8_vector_laplacian(level, m["elements"], fff, nn, vs, _dt(real))
    u = np.random.default_rng(level + vs).uniform(-1, 1, nn * vs)
    ref = orc.sshex8_vector_laplacian_apply(level, m["elements"], fff, vs, u)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    op.apply(x, y)  # accumulate semantics
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    y.fill_(9.0)
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    # every component takes the scalar operator's diagonal
    sop = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, _dt(real), level=level)
    ds = torch.zeros(nn, dtype=_dt(real), device="cuda")
    sop.diag(ds)
    d = torch.zeros_like(x)
    op.diag(d)
    # (the scalar diagonal is accumulated with REDs: two runs agree to rounding, not bit for bit)
    assert float(((d.reshape(nn, vs) - ds[:, None]).abs() / ds[:, None]).max()) < (1e-14 if real == "f64" else 1e-6)
    assert float((d.reshape(nn, vs) - d.reshape(nn, vs)[:, :1]).abs().max()) == 0.0


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [1, 2, 4, 8])
def test_sshex8_lumped_mass(oracles, real, level):
    """b200fem_sshex8_lumped_mass_create against affine_sshex8_mass_lumped (operators/sshex8/sshex8_mass.cpp:11-115) on a warped
    (non-affine) macro mesh; the total is the volume of the mesh."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    m = mesh.sshex8_box(level, 3, 2, 4)
    mpts = mesh.warp_points(m["m_points"])
    nn = m["n_nodes"]
    pts = np.zeros((3, nn), dtype=np.float32)  # coordinates by sshex8 node id: only the macro corners are read
    corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
    for k, c in enumerate(corners):
        pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
    op = lib.sshex8_lumped_mass(level, m["elements"], pts, nn, _dt(real))
    d = torch.zeros(nn, dtype=_dt(real), device="cuda")
    op.diag(d)
    ref = orc.sshex8_lumped_mass(level, m["elements"], pts, nn)
    assert _rel(d.cpu().numpy(), ref) < TOL[real]
    assert float(d.min()) > 0
    with pytest.raises(RuntimeError):
        op.apply(d, d)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_tet10_lumped_mass(oracles, real):
    """b200fem_mass_create(TET10).diag against tet10_assemble_lumped_mass (operators/tet10/tet10_mass.cpp:443-583): the row sums
    of the P2 mass matrix, negative on the corners. Bar 1e-10 in fp64 as for tet4 (the reference expands its determinant)."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    el, pts = mesh.box_tet4(6, 5, 7)
    el10, pts10 = mesh.tet4_to_tet10(el, mesh.warp_points(pts))
    nn = pts10.shape[1]
    op = lib.mass(lib.TET10, el10, pts10, nn, _dt(real))
    d = torch.zeros(nn, dtype=_dt(real), device="cuda")
    op.diag(d)
    ref = orc.tet10_lumped_mass(el10, pts10, nn)
    assert _rel(d.cpu().numpy(), ref) < (1e-10 if real == "f64" else 1e-4)
    assert float(d[: pts.shape[1]].max()) < 0 < float(d[pts.shape[1]:].min())
    with pytest.raises(RuntimeError):
        op.apply(d, d)


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", ["tet10", "macro_tet4"])
def test_p2_tet_from_points(oracles, real, family):
    """The 10-node tetrahedra from the points (b200fem_laplacian_create_from_points / b200fem_linear_elasticity_create_from_points
    with TET10 / MACRO_TET4) against tet10_* / macro_tet4_* _apply and _diag of the reference (geometry recomputed per element in
    the vector precision). tet10 elasticity fp64: 1e-11 (the reference's truncated quadrature points)."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    el4, pts4 = mesh.box_tet4(6, 7, 5)
    el, pts = mesh.tet4_to_tet10(el4, mesh.warp_points(pts4))
    nn = pts.shape[1]
    et = lib.TET10 if family == "tet10" else lib.MACRO_TET4
    dt = _dt(real)
    op = lib.laplacian_from_points(et, el, pts, nn, dt)
    u = np.random.default_rng(31).uniform(-1, 1, nn)
    x = torch.from_numpy(u).to(dt).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), orc.p2tet_laplacian_points(family, el, pts, u, nn)) < TOL[real]
    d = torch.zeros_like(x)
    op.diag(d)
    assert _rel(d.cpu().numpy(), orc.p2tet_laplacian_points(family, el, pts, None, nn, diag=True)) < TOL[real]
    mu, lam = 1.3, 0.7
    tol = (1e-11 if family == "tet10" else 1e-12) if real == "f64" else 1e-5
    le = lib.linear_elasticity_from_points(et, el, pts, mu, lam, nn, dt)
    u3 = np.random.default_rng(32).uniform(-1, 1, nn * 3)
    x3 = torch.from_numpy(u3).to(dt).cuda()
    y3 = torch.zeros_like(x3)
    le.apply(x3, y3)
    assert _rel(y3.cpu().numpy(), orc.p2tet_linear_elasticity_points(family, el, pts, mu, lam, u3, nn)) < tol
    y3.fill_(2.0)
    le.apply_overwrite(x3, y3)
    assert _rel(y3.cpu().numpy(), orc.p2tet_linear_elasticity_points(family, el, pts, mu, lam, u3, nn)) < tol
    d3 = torch.zeros_like(x3)
    le.diag(d3)
    assert _rel(d3.cpu().numpy(), orc.p2tet_linear_elasticity_points(family, el, pts, mu, lam, None, nn, diag=True)) < tol

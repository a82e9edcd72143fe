"""The oracle itself, on CPU: (1) the reference-compiled library (oracle/_ref, when present) and our plain-C
restatement (oracle/oracle_port.c) against the known answers the reference's own tests hold for this path, (2) the
restatement against the reference build on seeded inputs, (3) both against the committed vectors in tests/golden/
(generated from oracle/_ref by tests/golden/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
TOL = {"f64": 1e-12, "f32": 2e-5}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _kinds(oracles, real):
    return [k for k in ("reference", "port") if (k, real) in oracles]


def test_unit_tet_matrix_known_answer(oracles):
    """tests/ref/oracle.py of the reference: A_ij = grad_i . grad_j / 6 on the unit tet."""
    g = np.array([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
    A = g @ g.T / 6
    el = np.array([[0], [1], [2], [3]], dtype=np.int32)
    pts = np.array([[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    for kind in _kinds(oracles, "f64"):
        orc = oracles[(kind, "f64")]
        fff = orc.tet4_fff_fill(el, pts)
        for j in range(4):
            e = np.zeros(4)
            e[j] = 1
            assert np.allclose(orc.tet4_laplacian_apply_opt(el, fff, e), A[:, j], atol=1e-7), kind
        assert np.allclose(orc.tet4_laplacian_diag_opt(el, fff, 4), np.diag(A), atol=1e-7)


def test_mass_known_answers(oracles):
    """Mass oracles (reference build only): the unit tet's lumped mass is |T| / 4 = 1 / 24 per node, the consistent hex8 mass
    of the unit cube is the tensor product of the 1-D P1 mass matrix [[1/3, 1/6], [1/6, 1/3]], the mass of the constant is the
    lumped mass and sums to the volume."""
    if ("reference", "f64") not in oracles:
        pytest.skip("the mass routines are compiled from the reference only")
    orc = oracles[("reference", "f64")]
    el = np.array([[0], [1], [2], [3]], dtype=np.int32)
    pts = np.array([[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)
    assert np.allclose(orc.tet4_lumped_mass(el, pts, 4), 1.0 / 24.0, atol=1e-15)
    hel = np.arange(8, dtype=np.int32).reshape(8, 1)  # VTK corner order
    hpts = np.array([[0, 1, 1, 0, 0, 1, 1, 0], [0, 0, 1, 1, 0, 0, 1, 1], [0, 0, 0, 0, 1, 1, 1, 1]], dtype=np.float32)
    m1 = np.array([[1 / 3, 1 / 6], [1 / 6, 1 / 3]])
    M = np.zeros((8, 8))
    for i in range(8):
        for j in range(8):
            M[i, j] = np.prod([m1[int(hpts[d, i]), int(hpts[d, j])] for d in range(3)])
    for j in range(8):
        e = np.zeros(8)
        e[j] = 1
        assert np.allclose(orc.hex8_apply_mass(hel, hpts, e, 8), M[:, j], atol=1e-13)  # the 15-digit quadrature literals
    assert np.allclose(orc.hex8_lumped_mass(hel, hpts, 8), 1.0 / 8.0, atol=1e-13)
    from sfem_b200 import mesh

    bel, bpts = mesh.box_hex8(5, 4, 6)
    nn = bpts.shape[1]
    lumped = orc.hex8_lumped_mass(bel, bpts, nn)
    assert np.allclose(orc.hex8_apply_mass(bel, bpts, np.ones(nn), nn), lumped, atol=1e-14)
    assert abs(lumped.sum() - 1.0) < 1e-12


def test_stencil_known_answer(oracles):
    """Known answer in the spirit of operators/stencil/tests/sfem_StencilTest.cpp:143-218 (a quadratic field
    gives a constant at every interior stencil row): for u = x^2 + y^2 + z^2 on a unit-cube macro element the
    interior rows of the assembled level-L operator are  int grad u . grad phi_i = -int (Laplace u) phi_i = -6 h^3
    (h = 1/L) -- exact for trilinear elements on a uniform lattice."""
    from sfem_b200 import mesh

    for L in (4, 8):
        m = mesh.sshex8_box(L, 1, 1, 1)
        x, y, z = (m["points"][d].astype(np.float64) for d in range(3))
        u = x * x + y * y + z * z
        h = 1.0 / L
        for kind in _kinds(oracles, "f64"):
            orc = oracles[(kind, "f64")]
            fff = orc.hex8_fff_fill(m["m_elements"], m["m_points"])
            out = orc.sshex8_laplacian_apply(L, m["elements"], fff, u)
            interior = m["elements"][[mesh.sshex8_lidx(L, a, b, c) for a in range(1, L) for b in range(1, L) for c in range(1, L)], 0]
            assert np.allclose(out[interior], -6 * h * h * h, atol=1e-12), (kind, L)


def test_constants_and_symmetry(oracles):
    from sfem_b200 import mesh

    m = mesh.sshex8_box(4, 2, 2, 1)
    mp = mesh.shear_points(m["m_points"])
    nn = m["n_nodes"]
    rng = np.random.default_rng(0)
    a, b = rng.uniform(-1, 1, nn), rng.uniform(-1, 1, nn)
    for kind in _kinds(oracles, "f64"):
        orc = oracles[(kind, "f64")]
        fff = orc.hex8_fff_fill(m["m_elements"], mp)
        assert np.abs(orc.sshex8_laplacian_apply(4, m["elements"], fff, np.ones(nn))).max() < 1e-12
        Aa, Ab = orc.sshex8_laplacian_apply(4, m["elements"], fff, a), orc.sshex8_laplacian_apply(4, m["elements"], fff, b)
        assert abs(b @ Aa - a @ Ab) < 1e-11 * abs(b @ Aa)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_port_matches_reference_build(oracles, real):
    """Pins the restatement: every function of oracle_port.c against the reference's own code."""
    if ("reference", real) not in oracles or ("port", real) not in oracles:
        pytest.skip("needs both oracle libraries")
    from sfem_b200 import mesh

    ref, port = oracles[("reference", real)], oracles[("port", real)]
    tol = TOL[real]
    rng = np.random.default_rng(11)
    el, pts = mesh.box_tet4(6, 5, 4)
    pts = mesh.warp_points(pts)
    fr, fp = ref.tet4_fff_fill(el, pts), port.tet4_fff_fill(el, pts)
    assert _rel(fp, fr) < 2e-6
    u = rng.uniform(-1, 1, pts.shape[1])
    assert _rel(port.tet4_laplacian_apply_opt(el, fr, u), ref.tet4_laplacian_apply_opt(el, fr, u)) < tol
    assert _rel(port.tet4_laplacian_diag_opt(el, fr, pts.shape[1]), ref.tet4_laplacian_diag_opt(el, fr, pts.shape[1])) < tol

    el, pts = mesh.box_hex8(5, 6, 4)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    assert _rel(port.hex8_fff_fill(el, pts), ref.hex8_fff_fill(el, pts)) < 1e-6
    ar, dr = ref.hex8_adjugate_and_det_fill(el, pts)
    ap, dp = port.hex8_adjugate_and_det_fill(el, pts)
    assert _rel(ap, ar) < 1e-6 and _rel(dp, dr) < 1e-6
    fr = ref.hex8_fff_fill(el, pts)
    u = rng.uniform(-1, 1, nn)
    assert _rel(port.hex8_laplacian_apply_opt(el, fr, u), ref.hex8_laplacian_apply_opt(el, fr, u)) < tol
    u3 = rng.uniform(-1, 1, nn * 3)
    assert _rel(port.hex8_linear_elasticity_apply(el, ar, dr, 1.3, 0.7, u3, nn),
                ref.hex8_linear_elasticity_apply(el, ar, dr, 1.3, 0.7, u3, nn)) < tol

    for L in (2, 4):
        m = mesh.sshex8_box(L, 2, 1, 2)
        mp = mesh.shear_points(m["m_points"])
        fr = ref.hex8_fff_fill(m["m_elements"], mp)
        ar, dr = ref.hex8_adjugate_and_det_fill(m["m_elements"], mp)
        nn = m["n_nodes"]
        u = rng.uniform(-1, 1, nn)
        assert _rel(port.sshex8_laplacian_apply(L, m["elements"], fr, u), ref.sshex8_laplacian_apply(L, m["elements"], fr, u)) < tol
        u3 = rng.uniform(-1, 1, nn * 3)
        assert _rel(port.sshex8_linear_elasticity_apply(L, m["elements"], ar, dr, 1.3, 0.7, u3, nn),
                    ref.sshex8_linear_elasticity_apply(L, m["elements"], ar, dr, 1.3, 0.7, u3, nn)) < tol


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_golden_vectors(oracles, real):
    """Both oracle libraries reproduce the committed reference outputs."""
    for kind in _kinds(oracles, real):
        orc = oracles[(kind, real)]
        tol = TOL[real]
        g = np.load(os.path.join(GOLD, "tet4_laplacian.npz"))
        assert _rel(orc.tet4_laplacian_apply_opt(g["elements"], g["fff"], g["u"]), g[f"y_{real}"]) < tol, kind
        g = np.load(os.path.join(GOLD, "hex8.npz"))
        nn = g["points"].shape[1]
        assert _rel(orc.hex8_laplacian_apply_opt(g["elements"], g["fff"], g["u"]), g[f"y_lap_{real}"]) < tol, kind
        assert _rel(orc.hex8_linear_elasticity_apply(g["elements"], g["adjugate"], g["determinant"], float(g["mu"]), float(g["lam"]),
                                                     g["u3"], nn), g[f"y_le_{real}"]) < tol, kind
        for L in (2, 4, 8):
            g = np.load(os.path.join(GOLD, f"sshex8_L{L}.npz"))
            assert _rel(orc.sshex8_laplacian_apply(L, g["elements"], g["fff"], g["u"]), g[f"y_lap_{real}"]) < tol, (kind, L)
            if "y_le_f64" in g and real == "f64":
                nn = g["u"].size
                assert _rel(orc.sshex8_linear_elasticity_apply(L, g["elements"], g["adjugate"], g["determinant"], float(g["mu"]),
                                                               float(g["lam"]), g["u3"], nn), g["y_le_f64"]) < tol, (kind, L)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_port_restates_the_ten_node_laplacians(oracles, real):
    """oracle_port.c: tet10 / macro_tet4 Laplacian by textbook quadrature / sub-tet gradients against the reference's generated
    routines (apply and diag; the tet10 diagonal with the reference's factor 1/6) and against tests/golden/p2_tet.npz."""
    port = oracles.get(("port", real))
    if port is None:
        pytest.skip("port oracle not built")
    g = np.load(os.path.join(GOLD, "p2_tet.npz"))
    el, fff, nn = g["elements"], g["fff"], g["points"].shape[1]
    tol = TOL[real]
    for fam in ("tet10", "macro_tet4"):
        assert _rel(port.p2tet_laplacian_apply_opt(fam, el, fff, g["u"]), g[f"{fam}_lap_{real}"]) < tol, fam
        if real == "f64":
            assert _rel(port.p2tet_laplacian_diag_opt(fam, el, fff, nn), g[f"{fam}_lap_diag_f64"]) < tol, fam
    ref = oracles.get(("reference", real))
    if ref is not None:
        from sfem_b200 import mesh

        el4, pts4 = mesh.box_tet4(5, 4, 6)
        el10, pts10 = mesh.tet4_to_tet10(el4, mesh.warp_points(pts4))
        f = ref.tet4_fff_fill(el10[:4], pts10)
        u = np.random.default_rng(3).uniform(-1, 1, pts10.shape[1])
        for fam in ("tet10", "macro_tet4"):
            assert _rel(port.p2tet_laplacian_apply_opt(fam, el10, f, u), ref.p2tet_laplacian_apply_opt(fam, el10, f, u)) < tol, fam


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", ["tet4", "hex8", "tet10", "macro_tet4"])
def test_port_packed_laplacian_pack_loop(oracles, real, family):
    """The restated pack loop of PackedLaplacianApply::apply and ::apply_two_pass (frontend/ops/sfem_PackedLaplacian.cpp:111-262, a
    file that needs smesh and cannot be compiled here) gives the unpacked operator on the expanded connectivity -- the reference's
    own *_laplacian_apply_opt where oracle/_ref exists -- for both forms."""
    from sfem_b200 import lib, mesh

    port = oracles.get(("port", real))
    if port is None:
        pytest.skip("port oracle not built")
    ref = oracles.get(("reference", real)) or port
    if family == "hex8":
        el, pts = mesh.box_hex8(7, 6, 5)
        pts = mesh.shear_points(pts)
        et, epp = 8, 32
    else:
        el, pts = mesh.box_tet4(5, 6, 4)
        pts = mesh.warp_points(pts)
        et, epp = 4, 100
        if family != "tet4":
            el, pts = mesh.tet4_to_tet10(el, pts)
            et = 10 if family == "tet10" else 100
    nn = pts.shape[1]
    pm = lib.PackedMesh(el, nn, epp)  # host code of libb200fem.so: needs no GPU
    gl = pm.global_elements()
    pp = np.ascontiguousarray(pts[:, pm.new_to_old])
    if family == "hex8":
        fff = ref.hex8_fff_fill(gl, pp)
        expect = lambda u: ref.hex8_laplacian_apply_opt(gl, fff, u)
    elif family == "tet4":
        fff = ref.tet4_fff_fill(gl, pp)
        expect = lambda u: ref.tet4_laplacian_apply_opt(gl, fff, u)
    else:
        fff = ref.tet4_fff_fill(gl[:4], pp)
        expect = lambda u: ref.p2tet_laplacian_apply_opt(family, gl, fff, u)
    u = np.random.default_rng(9).uniform(-1, 1, nn)
    want = expect(u)
    assert _rel(port.packed_laplacian_apply(et, pm, fff, u), want) < TOL[real]
    assert _rel(port.packed_laplacian_apply(et, pm, fff, u, two_pass=True), want) < TOL[real]


@pytest.mark.parametrize("family", ["tet10", "macro_tet4"])
def test_port_ten_node_elasticity_and_the_reference_quadrature_points(oracles, family):
    """oracle_port.c: tet10 / macro_tet4 linear elasticity from the textbook weak form (exact 4-point rule / constant-strain sub-tets)
    against the reference's routines. A second, independent witness of the quirk DESIGN.md section 2 states: the reference's tet10
    elasticity (face-centroid points written as 0.333333333333) is ~1.5e-12 from the exact integral, macro_tet4 is at rounding."""
    port, ref = oracles.get(("port", "f64")), oracles.get(("reference", "f64"))
    if port is None or ref is None:
        pytest.skip("needs both oracle builds")
    g = np.load(os.path.join(GOLD, "p2_tet.npz"))
    el, adj, det = g["elements"], g["adjugate"], g["determinant"]
    mine = port.p2tet_linear_elasticity_apply_opt(family, el, adj, det, float(g["mu"]), float(g["lam"]), g["u3"])
    theirs = ref.p2tet_linear_elasticity_apply_opt(family, el, adj, det, float(g["mu"]), float(g["lam"]), g["u3"])
    assert _rel(theirs, g[f"{family}_le_f64"]) < 1e-14
    err = _rel(mine, theirs)
    if family == "tet10":
        assert 1e-13 < err < 1e-11, err  # the reference's truncated quadrature points, not rounding
    else:
        assert err < 1e-13, err

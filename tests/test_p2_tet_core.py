"""The per-element arithmetic of the tet10 / macro_tet4 CUDA kernels (sfem_b200/csrc/p2_tet_local.cuh: corner-gradient form of
the P2 operators instead of the reference's 8-point quadrature, table-driven sub-tets instead of its generated sub-FFF /
sub-adjugate routines) run sequentially on the host -- the same header, compiled with g++ -- against the reference's own
tet10_* / macro_tet4_* routines compiled into the oracle. Bars: 1e-12 rel. L2 (fp64), 1e-5 (fp32)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = ["tet10", "macro_tet4"]


@pytest.fixture(scope="module")
def core_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("p2core") / "libp2_tet_core_host.so"
    src = os.path.join(ROOT, "tests", "helpers", "p2_tet_core_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "sfem_b200", "csrc"),
                    "-o", str(out), src], check=True)
    return C.CDLL(str(out))


def _ref(oracles, real):
    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    return orc


def _mesh():
    from sfem_b200 import mesh

    el, pts = mesh.box_tet4(5, 4, 6)
    pts = mesh.warp_points(pts)
    el10, pts10 = mesh.tet4_to_tet10(el, pts)
    return el, pts, el10, pts10


def tet4_adjugate_and_det(el, pts):
    """numpy restatement of tet4_adjugate_and_det (operators/tet4/tet4_inline_cpu.hpp:285-380) in fp32 storage:
    J[r][c] = x_r(v_{c+1}) - x_r(v_0), adjugate = det J^-1 (row-major AoS [ne*9]), determinant [ne]."""
    p = pts.astype(np.float64)
    J = np.stack([p[:, el[c + 1]] - p[:, el[0]] for c in range(3)], axis=1)  # [r][c][ne]
    J = np.moveaxis(J, 2, 0)  # [ne][r][c]
    det = np.linalg.det(J)
    adj = np.linalg.inv(J) * det[:, None, None]
    return adj.reshape(-1).astype(np.float32), det.astype(np.float32)


def test_macro_tet4_table_is_the_inverse_of_the_subtet_maps(core_lib):
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [.5, 0, 0], [.5, .5, 0], [0, .5, 0], [0, 0, .5], [.5, 0, .5], [0, .5, .5]])
    for s in range(8):
        nodes = (C.c_int * 4)()
        hb = (C.c_int * 9)()
        core_lib.p2_tet_macro_table(s, nodes, hb)
        n = list(nodes)
        A = np.stack([P[n[1]] - P[n[0]], P[n[2]] - P[n[0]], P[n[3]] - P[n[0]]], axis=1)
        assert abs(np.linalg.det(A) - 0.125) < 1e-15
        assert np.array_equal(np.linalg.inv(A).round().astype(int), 2 * np.array(list(hb)).reshape(3, 3))


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", FAMILIES)
def test_laplacian_core_matches_reference(core_lib, oracles, real, family):
    from oracle.ref import _soa_ptrs

    orc = _ref(oracles, real)
    el, pts, el10, pts10 = _mesh()
    nn = pts10.shape[1]
    fff = orc.tet4_fff_fill(el, pts)
    fn = getattr(core_lib, "p2_tet_laplacian_" + real)
    macro = int(family == "macro_tet4")
    tol = 1e-12 if real == "f64" else 1e-5
    el10c = np.ascontiguousarray(el10)
    for seed in (0, 1):
        u = np.random.default_rng(seed).uniform(-1, 1, nn).astype(orc.rt)
        out = np.zeros(nn, dtype=orc.rt)
        fn(macro, 0, 1, C.c_ssize_t(el10c.shape[1]), _soa_ptrs(el10c, C.c_int32), fff.ctypes.data_as(C.POINTER(C.c_float)),
           u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        ref = orc.p2tet_laplacian_apply_opt(family, el10, fff, u)
        assert np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref) < tol
    d = np.zeros(nn, dtype=orc.rt)
    fn(macro, 1, 1, C.c_ssize_t(el10c.shape[1]), _soa_ptrs(el10c, C.c_int32), fff.ctypes.data_as(C.POINTER(C.c_float)), None,
       d.ctypes.data_as(C.c_void_p))
    ref = orc.p2tet_laplacian_diag_opt(family, el10, fff, nn)
    assert np.linalg.norm(d.astype(np.float64) - ref) / np.linalg.norm(ref) < tol


def test_tet10_reference_diag_is_a_sixth_of_its_apply(core_lib, oracles):
    """Pins the reference behaviour the kernel reproduces by default: tet10_laplacian_diag_opt returns 1/6 of the diagonal of
    tet10_laplacian_apply_opt (the generated diagonal misses the reference volume the FFF and the apply carry)."""
    from oracle.ref import _soa_ptrs

    orc = _ref(oracles, "f64")
    el, pts, el10, pts10 = _mesh()
    nn = pts10.shape[1]
    fff = orc.tet4_fff_fill(el, pts)
    ref_diag = orc.p2tet_laplacian_diag_opt("tet10", el10, fff, nn)
    idx = np.arange(0, nn, 53)
    true = np.empty(idx.size)
    for j, i in enumerate(idx):
        e = np.zeros(nn)
        e[i] = 1
        true[j] = orc.p2tet_laplacian_apply_opt("tet10", el10, fff, e)[i]
    assert np.allclose(6 * ref_diag[idx], true, rtol=1e-12)
    d = np.zeros(nn)
    el10c = np.ascontiguousarray(el10)
    core_lib.p2_tet_laplacian_f64(0, 1, 0, C.c_ssize_t(el10c.shape[1]), _soa_ptrs(el10c, C.c_int32),
                                  fff.ctypes.data_as(C.POINTER(C.c_float)), None, d.ctypes.data_as(C.c_void_p))
    assert np.allclose(d[idx], true, rtol=1e-12)


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", FAMILIES)
def test_elasticity_core_matches_reference(core_lib, oracles, real, family):
    from oracle.ref import _soa_ptrs

    orc = _ref(oracles, real)
    el, pts, el10, pts10 = _mesh()
    nn = pts10.shape[1]
    adj, det = tet4_adjugate_and_det(el, pts)
    mu, lam = 1.3, 0.7
    fn = getattr(core_lib, "p2_tet_elasticity_" + real)
    macro = int(family == "macro_tet4")
    # tet10: the reference's face-centroid quadrature points are written as 0.333333333333 (12 digits,
    # operators/tet10/tet10_linear_elasticity.cpp:10-17), which puts ITS result 1.5e-12 from the exact integral; this code
    # (checked against a long-double evaluation: 1.7e-16) integrates exactly, hence the 1e-11 bar for that family
    tol = (1e-11 if family == "tet10" else 1e-12) if real == "f64" else 1e-5
    el10c = np.ascontiguousarray(el10)
    u = np.random.default_rng(5).uniform(-1, 1, nn * 3).astype(orc.rt)
    out = np.zeros(nn * 3, dtype=orc.rt)
    args = (C.c_ssize_t(el10c.shape[1]), _soa_ptrs(el10c, C.c_int32), adj.ctypes.data_as(C.POINTER(C.c_float)),
            det.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(mu), C.c_double(lam))
    fn(macro, 0, *args, u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    ref = orc.p2tet_linear_elasticity_apply_opt(family, el10, adj, det, mu, lam, u)
    assert np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref) < tol
    d = np.zeros(nn * 3, dtype=orc.rt)
    fn(macro, 1, *args, None, d.ctypes.data_as(C.c_void_p))
    ref = orc.p2tet_linear_elasticity_diag_opt(family, el10, adj, det, mu, lam, nn)
    assert np.linalg.norm(d.astype(np.float64) - ref) / np.linalg.norm(ref) < tol


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", FAMILIES)
def test_core_and_reference_reproduce_the_golden_vectors(core_lib, oracles, real, family):
    """tests/golden/p2_tet.npz (written by `tests/golden/make_golden.py p2_tet` from oracle/_ref): the host run of the kernels'
    header reproduces it, and so does the reference build where it is present."""
    from oracle.ref import _soa_ptrs

    g = np.load(os.path.join(ROOT, "tests", "golden", "p2_tet.npz"))
    el = np.ascontiguousarray(g["elements"])
    nn = g["points"].shape[1]
    rt = np.float64 if real == "f64" else np.float32
    macro = int(family == "macro_tet4")
    tol = 1e-12 if real == "f64" else 1e-5
    le_tol = (1e-11 if family == "tet10" else 1e-12) if real == "f64" else 1e-5
    rel = lambda a, b: np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b)
    fff, adj, det = (np.ascontiguousarray(g[k]) for k in ("fff", "adjugate", "determinant"))
    pf = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    u = g["u"].astype(rt)
    out = np.zeros(nn, dtype=rt)
    getattr(core_lib, "p2_tet_laplacian_" + real)(macro, 0, 1, C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf(fff),
                                                  u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rel(out, g[f"{family}_lap_{real}"]) < tol
    u3 = g["u3"].astype(rt)
    out3 = np.zeros(nn * 3, dtype=rt)
    getattr(core_lib, "p2_tet_elasticity_" + real)(macro, 0, C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf(adj), pf(det),
                                                   C.c_double(float(g["mu"])), C.c_double(float(g["lam"])),
                                                   u3.ctypes.data_as(C.c_void_p), out3.ctypes.data_as(C.c_void_p))
    assert rel(out3, g[f"{family}_le_{real}"]) < le_tol
    orc = oracles.get(("reference", real))
    if orc is not None:
        assert rel(orc.p2tet_laplacian_apply_opt(family, el, fff, g["u"]), g[f"{family}_lap_{real}"]) < tol
        assert rel(orc.p2tet_linear_elasticity_apply_opt(family, el, adj, det, float(g["mu"]), float(g["lam"]), g["u3"]),
                   g[f"{family}_le_{real}"]) < tol


def test_tet10_is_exact_on_harmonic_quadratics(core_lib):
    """Known answer that does not go through the reference: a quadratic field lies in the P2 space (straight-edged elements,
    mid-edge nodes), so for a HARMONIC quadratic u the Galerkin rows of the interior nodes vanish, (K u)_i = -int phi_i lap(u) = 0,
    whatever the mesh; the P1-iso-P2 operator (macro_tet4) has that property for linear fields only."""
    from oracle.ref import _soa_ptrs
    from sfem_b200 import mesh

    el4, pts4 = mesh.box_tet4(5, 6, 4)
    el, pts = mesh.tet4_to_tet10(el4, pts4)
    nn = pts.shape[1]
    p = pts.astype(np.float64)
    # FFF of the corner tets in fp64, stored fp32 like the plan's cache
    J = np.moveaxis(np.stack([p[:, el[c + 1]] - p[:, el[0]] for c in range(3)], axis=1), 2, 0)
    det = np.linalg.det(J)
    Ji = np.linalg.inv(J)
    F = np.einsum("eij,ekj->eik", Ji, Ji) * (det / 6.0)[:, None, None]
    fff = np.ascontiguousarray(np.stack([F[:, 0, 0], F[:, 0, 1], F[:, 0, 2], F[:, 1, 1], F[:, 1, 2], F[:, 2, 2]], axis=1).reshape(-1),
                               dtype=np.float32)
    interior = np.all((p > 1e-6) & (p < 1 - 1e-6), axis=0)
    assert interior.sum() > 50
    elc = np.ascontiguousarray(el)

    def apply(macro, u):
        out = np.zeros(nn)
        core_lib.p2_tet_laplacian_f64(macro, 0, 1, C.c_ssize_t(elc.shape[1]), _soa_ptrs(elc, C.c_int32),
                                      fff.ctypes.data_as(C.POINTER(C.c_float)), u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    x, y, z = p
    harmonic = x * x - 0.5 * y * y - 0.5 * z * z + 0.3 * x * y - 0.2 * y * z + x - 2 * z + 1
    linear = 0.7 * x - 1.3 * y + 0.4 * z + 2
    scale = np.abs(apply(0, x * x)).max()
    # fp32 geometry: the rows vanish to the rounding of the stored FFF (1e-7 relative), far below the O(1) rows of x^2
    assert np.abs(apply(0, harmonic)[interior]).max() < 5e-6 * scale
    assert np.abs(apply(0, x * x)[interior]).max() > 0.01 * scale  # lap(x^2) = 2: not harmonic, rows do not vanish
    assert np.abs(apply(1, linear)[interior]).max() < 5e-6 * scale
    assert np.abs(apply(1, harmonic)[interior]).max() > 1e-3 * scale  # P1-iso-P2 is not exact on quadratics


def test_tet4_block_diag_assembly_and_the_reference_factor_six(core_lib, oracles):
    """tet4_linear_elasticity_block_diag_sym (operators/tet4/tet4_linear_elasticity.cpp:360-440) calls its point kernel with weight 1
    instead of 1/6: its blocks are 6 x the node blocks of tet4_linear_elasticity_apply / _diag. Pinned here, and the assembly loop of
    tet4_le_block_diag_kernel (restated on the host over the same local operator) reproduces it with scale 6."""
    from oracle.ref import _soa_ptrs
    from sfem_b200 import mesh

    orc = _ref(oracles, "f64")
    el, pts = mesh.box_tet4(5, 4, 6)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    mu, lam = 1.3, 0.7
    ref = orc.tet4_linear_elasticity_block_diag_sym(el, pts, mu, lam, nn)
    d = orc.tet4_linear_elasticity_diag_points(el, pts, mu, lam, nn)
    assert np.allclose(ref.reshape(nn, 6)[:, [0, 3, 5]], 6 * d.reshape(nn, 3), rtol=1e-12)
    adj, det = tet4_adjugate_and_det(el, pts)
    elc = np.ascontiguousarray(el)
    out = np.zeros(nn * 6)
    core_lib.tet4_le_block_diag_f64(C.c_ssize_t(elc.shape[1]), _soa_ptrs(elc, C.c_int32), adj.ctypes.data_as(C.POINTER(C.c_float)),
                                    det.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(mu), C.c_double(lam), C.c_double(6.0),
                                    out.ctypes.data_as(C.c_void_p))
    # fp32 adjugate here against the reference's fp64 one from the points: 1e-6
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-6
    # off-diagonal entries against unit applies of the reference
    i = 37
    for b in range(3):
        e = np.zeros(nn * 3)
        e[i * 3 + b] = 1
        row = orc.tet4_linear_elasticity_apply_points(el, pts, mu, lam, e, nn).reshape(nn, 3)[i]
        blk = ref.reshape(nn, 6)[i]
        full = np.array([[blk[0], blk[1], blk[2]], [blk[1], blk[3], blk[4]], [blk[2], blk[4], blk[5]]])
        assert np.allclose(6 * row, full[b], rtol=1e-11)


@pytest.mark.parametrize("family", FAMILIES)
def test_elasticity_rigid_modes_and_patch_test(core_lib, family):
    """Known answers that do not go through the reference: rigid body motions (translation + infinitesimal rotation) lie in the null
    space of the elasticity operator on every row, and a constant-strain field u = E x (in both the P2 and the P1-iso-P2 space)
    produces a divergence-free stress, so the rows of the interior nodes vanish (patch test)."""
    from oracle.ref import _soa_ptrs
    from sfem_b200 import mesh

    el4, pts4 = mesh.box_tet4(4, 5, 4)
    el, pts = mesh.tet4_to_tet10(el4, pts4)
    nn = pts.shape[1]
    adj, det = tet4_adjugate_and_det(el4, pts4)
    elc = np.ascontiguousarray(el)
    macro = int(family == "macro_tet4")

    def apply(u):
        out = np.zeros(nn * 3)
        core_lib.p2_tet_elasticity_f64(macro, 0, C.c_ssize_t(elc.shape[1]), _soa_ptrs(elc, C.c_int32), adj.ctypes.data_as(C.POINTER(C.c_float)),
                                       det.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(1.3), C.c_double(0.7),
                                       np.ascontiguousarray(u).ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out.reshape(nn, 3)

    p = pts.astype(np.float64).T  # [nn][3]
    stretch = p * np.array([1.0, -0.5, 0.25])  # reference magnitude: a non-trivial strain
    scale = np.abs(apply(stretch.reshape(-1))).max()
    assert scale > 0.01
    omega = np.array([0.3, -0.7, 0.5])
    rigid = np.array([0.2, -1.0, 0.4]) + np.cross(np.broadcast_to(omega, p.shape), p)
    assert np.abs(apply(rigid.reshape(-1))).max() < 1e-6 * scale  # fp32 adjugates: exact up to their rounding (dyadic box: exact)
    E = np.array([[0.5, 0.2, -0.1], [0.3, -0.4, 0.6], [0.0, 0.7, 0.1]])
    interior = np.all((p > 1e-6) & (p < 1 - 1e-6), axis=1)
    assert interior.sum() > 20
    r = apply((p @ E.T).reshape(-1))
    assert np.abs(r[interior]).max() < 1e-6 * scale
    assert np.abs(r[~interior]).max() > 1e-2 * scale  # boundary rows carry the traction


def test_tet10_elasticity_core_equals_the_exact_restatement(core_lib, oracles):
    """The corner-gradient form of the kernels' header against the textbook restatement of oracle_port.c (exact 4-point rule): both
    integrate exactly and agree to rounding, where the reference (truncated quadrature points) sits 1.5e-12 away."""
    from oracle.ref import _soa_ptrs

    port = oracles.get(("port", "f64"))
    if port is None:
        pytest.skip("port oracle not built")
    g = np.load(os.path.join(ROOT, "tests", "golden", "p2_tet.npz"))
    el = np.ascontiguousarray(g["elements"])
    nn = g["points"].shape[1]
    adj, det = np.ascontiguousarray(g["adjugate"]), np.ascontiguousarray(g["determinant"])
    u3 = np.ascontiguousarray(g["u3"])
    out = np.zeros(nn * 3)
    core_lib.p2_tet_elasticity_f64(0, 0, C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), adj.ctypes.data_as(C.POINTER(C.c_float)),
                                   det.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(float(g["mu"])), C.c_double(float(g["lam"])),
                                   u3.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    exact = port.p2tet_linear_elasticity_apply_opt("tet10", el, adj, det, float(g["mu"]), float(g["lam"]), u3)
    assert np.linalg.norm(out - exact) / np.linalg.norm(exact) < 1e-14
    assert np.linalg.norm(out - g["tet10_le_f64"]) / np.linalg.norm(exact) > 1e-13

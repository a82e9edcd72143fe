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

"""The drop-in boundary, compiled: sfem_b200/adapter/sfem_b200_ops.cpp (sfem::Op subclasses, register_b200_ops, B200BLAS) built
against the reference's own frontend/ops/sfem_Op.hpp, sfem_OpFactory.hpp, sfem_MultiDomainOp.{hpp,cpp} (oracle/Makefile target
`adapter`, smesh stand-ins in oracle/ref_shim/frontend_stub) and driven the way the reference's drivers do:
sfem::create_op(space, "Laplacian" | "LinearElasticity", EXECUTION_SPACE_DEVICE) -> initialize() -> apply(nullptr, h, out).
The factory must hand back OUR operators under the reference's keys ("gpu:" / "ss:gpu:" prefixes added by the factory),
and their results must equal the oracle's."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libsfem_ref_adapter_f64.so")


def _lib():
    if not os.path.exists(LIB):
        pytest.skip("adapter harness not built (make -C oracle adapter; needs the reference checkout)")
    h = C.CDLL(LIB)
    vp, i32, ssz = C.c_void_p, C.c_int, C.c_ssize_t
    pp_i = C.POINTER(C.POINTER(C.c_int32))
    pp_f = C.POINTER(C.POINTER(C.c_float))
    sig = {"adapter_space_create": (vp, [i32, i32, i32, ssz, ssz, pp_i, pp_f, i32]), "adapter_space_destroy": (None, [vp]),
           "adapter_create_op": (vp, [vp, C.c_char_p]), "adapter_op_destroy": (None, [vp]), "adapter_op_name": (C.c_char_p, [vp]),
           "adapter_op_execution_space": (i32, [vp]), "adapter_op_n_dofs": (ssz, [vp]), "adapter_op_apply": (i32, [vp, vp, vp]),
           "adapter_op_gradient": (i32, [vp, vp, vp]), "adapter_op_hessian_diag": (i32, [vp, vp]),
           "adapter_op_hessian_block_diag_sym": (i32, [vp, vp]), "adapter_op_clone": (vp, [vp]), "adapter_op_derefine": (vp, [vp, vp]),
           "adapter_blas_axpby_dot": (C.c_double, [ssz, C.c_double, vp, C.c_double, vp])}
    for n, (r, a) in sig.items():
        getattr(h, n).restype, getattr(h, n).argtypes = r, a
    return h


def test_adapter_library_exports():
    """CPU: the adapter built against the reference headers loads and exports its entry points."""
    h = _lib()
    for n in ("adapter_space_create", "adapter_create_op", "adapter_op_apply", "adapter_op_hessian_diag"):
        assert hasattr(h, n)


def _space(h, elem_type, level, el, pts, bs):
    from sfem_b200.lib import _soa_ptrs

    el = np.ascontiguousarray(el, dtype=np.int32)
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    return h.adapter_space_create(elem_type, level, el.shape[0], el.shape[1], pts.shape[1], _soa_ptrs(el, C.c_int32),
                                  _soa_ptrs(pts, C.c_float), bs)


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["tet4", "tet10", "macro_tet4", "hex8", "sshex8"])
def test_factory_creates_our_laplacian(oracles, case):
    import torch

    from sfem_b200 import mesh

    h = _lib()
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    level = 0
    if case == "tet4":
        el, pts = mesh.box_tet4(7, 6, 8)
        pts = mesh.warp_points(pts)
        nn = pts.shape[1]
        u = np.random.default_rng(1).uniform(-1, 1, nn)
        ref = orc.tet4_laplacian_apply_opt(el, orc.tet4_fff_fill(el, pts), u)
        sp = _space(h, 4, 0, el, pts, 1)
        key = b"gpu:Laplacian"
    elif case in ("tet10", "macro_tet4"):
        # smesh::TET10 = 10, smesh::MACRO_TET4 = 100: the 10-node families, geometry of the 4 corners
        orc = oracles.get(("reference", "f64"))
        if orc is None:
            pytest.skip("oracle/_ref not built")
        el4, pts4 = mesh.box_tet4(6, 5, 7)
        el, pts = mesh.tet4_to_tet10(el4, mesh.warp_points(pts4))
        nn = pts.shape[1]
        u = np.random.default_rng(6).uniform(-1, 1, nn)
        ref = orc.p2tet_laplacian_apply_opt(case, el, orc.tet4_fff_fill(el[:4], pts), u)
        sp = _space(h, 10 if case == "tet10" else 100, 0, el, pts, 1)
        key = b"gpu:Laplacian"
    elif case == "hex8":
        el, pts = mesh.box_hex8(6, 7, 5)
        pts = mesh.shear_points(pts)
        nn = pts.shape[1]
        u = np.random.default_rng(2).uniform(-1, 1, nn)
        ref = orc.hex8_laplacian_apply_opt(el, orc.hex8_fff_fill(el, pts), u)
        sp = _space(h, 8, 0, el, pts, 1)
        key = b"gpu:Laplacian"
    else:
        level = 4
        m = mesh.sshex8_box(level, 3, 2, 2)
        mpts = mesh.shear_points(m["m_points"])
        nn = m["n_nodes"]
        u = np.random.default_rng(3).uniform(-1, 1, nn)
        ref = orc.sshex8_laplacian_apply(level, m["elements"], orc.hex8_fff_fill(m["m_elements"], mpts), u)
        # node coordinates by sshex8 node id: only the macro corners are read (as in the reference's geometry caches)
        pts = np.zeros((3, nn), dtype=np.float32)
        corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
        for k, c in enumerate(corners):
            pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
        el = m["elements"]
        sp = _space(h, 800, level, el, pts, 1)
        key = b"ss:gpu:Laplacian"
    op = h.adapter_create_op(sp, b"Laplacian")  # the reference's call: name without prefixes, execution space DEVICE
    assert op, "Factory::create_op returned no operator"
    assert h.adapter_op_name(op) == key and h.adapter_op_execution_space(op) == 1 and h.adapter_op_n_dofs(op) == nn
    x = torch.from_numpy(u).cuda()
    y = torch.zeros_like(x)
    assert h.adapter_op_apply(op, x.data_ptr(), y.data_ptr()) == 0
    # fp32 geometry computed by b200fem_fff_fill (fp64 arithmetic -> fp32) against the reference's fill: 1e-6; the kernels
    # themselves are pinned at 1e-12 with identical FFF arrays in tests/test_gpu_operators.py
    assert _rel(y.cpu().numpy(), ref) < 1e-6
    assert h.adapter_op_apply(op, x.data_ptr(), y.data_ptr()) == 0  # Op::apply accumulates
    assert _rel(y.cpu().numpy(), 2 * ref) < 1e-6
    g = torch.zeros_like(x)
    assert h.adapter_op_gradient(op, x.data_ptr(), g.data_ptr()) == 0 and _rel(g.cpu().numpy(), ref) < 1e-6
    d = torch.zeros_like(x)
    assert h.adapter_op_hessian_diag(op, d.data_ptr()) == 0 and float(d.min()) > 0
    c = h.adapter_op_clone(op)
    y.zero_()
    assert c and h.adapter_op_apply(c, x.data_ptr(), y.data_ptr()) == 0 and _rel(y.cpu().numpy(), ref) < 1e-6
    # the adapter's sfem::BLAS<real_t>: y = 2 x - y; (x, y)
    y2 = torch.from_numpy(ref.copy()).cuda()
    dot = h.adapter_blas_axpby_dot(nn, 2.0, x.data_ptr(), -1.0, y2.data_ptr())
    assert abs(dot - float(u @ (2 * u - ref))) < 1e-9 * abs(dot)
    h.adapter_op_destroy(c)
    h.adapter_op_destroy(op)
    h.adapter_space_destroy(sp)


@pytest.mark.gpu
def test_factory_creates_our_mass(oracles):
    """sfem::create_op(space, "Mass" | "LumpedMass", DEVICE): Mass::apply on a hex8 block, LumpedMass::hessian_diag on hex8 and
    tet4 blocks, against the reference's own hex8_mass.cpp / tet4_mass.cpp."""
    import torch

    from sfem_b200 import mesh

    h = _lib()
    if ("reference", "f64") not in oracles:
        pytest.skip("the mass routines exist in the reference build of the oracle only")
    orc = oracles[("reference", "f64")]
    el, pts = mesh.box_hex8(6, 7, 5)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    u = np.random.default_rng(4).uniform(-1, 1, nn)
    sp = _space(h, 8, 0, el, pts, 1)
    op = h.adapter_create_op(sp, b"Mass")
    assert op and h.adapter_op_name(op) == b"gpu:Mass" and h.adapter_op_execution_space(op) == 1
    x = torch.from_numpy(u).cuda()
    y = torch.zeros_like(x)
    assert h.adapter_op_apply(op, x.data_ptr(), y.data_ptr()) == 0
    assert _rel(y.cpu().numpy(), orc.hex8_apply_mass(el, pts, u, nn)) < 1e-12
    lm = h.adapter_create_op(sp, b"LumpedMass")
    d = torch.zeros_like(x)
    assert lm and h.adapter_op_hessian_diag(lm, d.data_ptr()) == 0
    assert _rel(d.cpu().numpy(), orc.hex8_lumped_mass(el, pts, nn)) < 1e-12
    tel, tpts = mesh.box_tet4(5, 6, 4)
    tn = tpts.shape[1]
    tsp = _space(h, 4, 0, tel, tpts, 1)
    tlm = h.adapter_create_op(tsp, b"LumpedMass")
    td = torch.zeros(tn, dtype=torch.float64, device="cuda")
    assert tlm and h.adapter_op_hessian_diag(tlm, td.data_ptr()) == 0
    assert _rel(td.cpu().numpy(), orc.tet4_lumped_mass(tel, tpts, tn)) < 1e-10
    for o in (op, lm, tlm):
        h.adapter_op_destroy(o)
    h.adapter_space_destroy(sp)
    h.adapter_space_destroy(tsp)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["hex8", "sshex8", "tet4", "tet10", "macro_tet4"])
def test_factory_creates_our_linear_elasticity(oracles, case):
    import torch

    from sfem_b200 import mesh

    h = _lib()
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    if case == "hex8":
        el, pts = mesh.box_hex8(5, 6, 4, 4.0, 5.0, 3.0)
        pts = mesh.shear_points(pts, 0.25, 0.5, 0.125)  # dyadic: both sides derive the same fp32 adjugates
        nn = pts.shape[1]
        adj, det = orc.hex8_adjugate_and_det_fill(el, pts)
        u = np.random.default_rng(4).uniform(-1, 1, nn * 3)
        ref = orc.hex8_linear_elasticity_apply(el, adj, det, 1.0, 1.0, u, nn)
        sp = _space(h, 8, 0, el, pts, 3)
        key = b"gpu:LinearElasticity"
    elif case in ("tet4", "tet10", "macro_tet4"):
        from sfem_b200 import lib

        orc = oracles.get(("reference", "f64"))
        if orc is None:
            pytest.skip("oracle/_ref not built")
        el4, pts4 = mesh.box_tet4(5, 6, 4, 4.0, 5.0, 3.0)
        pts4 = mesh.shear_points(pts4, 0.25, 0.5, 0.125)
        nn4 = pts4.shape[1]
        if case == "tet4":
            el, pts, nn = el4, pts4, nn4
            u = np.random.default_rng(7).uniform(-1, 1, nn * 3)
            ref = orc.tet4_linear_elasticity_apply_points(el, pts, 1.0, 1.0, u, nn)
        else:
            el, pts = mesh.tet4_to_tet10(el4, pts4)
            nn = pts.shape[1]
            adj, det = lib.adjugate_and_det_fill(lib.TET4, el4, pts4)  # dyadic mesh: exact in fp32
            u = np.random.default_rng(7).uniform(-1, 1, nn * 3)
            ref = orc.p2tet_linear_elasticity_apply_opt(case, el, adj, det, 1.0, 1.0, u)
        sp = _space(h, {"tet4": 4, "tet10": 10, "macro_tet4": 100}[case], 0, el, pts, 3)
        key = b"gpu:LinearElasticity"
    else:
        level = 4
        m = mesh.sshex8_box(level, 2, 3, 2, 2.0, 3.0, 2.0)
        mpts = mesh.shear_points(m["m_points"], 0.25, 0.5, 0.125)
        nn = m["n_nodes"]
        adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
        u = np.random.default_rng(5).uniform(-1, 1, nn * 3)
        ref = orc.sshex8_linear_elasticity_apply(level, m["elements"], adj, det, 1.0, 1.0, u, nn)
        pts = np.zeros((3, nn), dtype=np.float32)
        corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
        for k, c in enumerate(corners):
            pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
        sp = _space(h, 800, level, m["elements"], pts, 3)
        key = b"ss:gpu:LinearElasticity"
    op = h.adapter_create_op(sp, b"LinearElasticity")  # mu = lambda = 1: the reference's defaults
    assert op and h.adapter_op_name(op) == key and h.adapter_op_n_dofs(op) == nn * 3
    x = torch.from_numpy(u).cuda()
    y = torch.zeros_like(x)
    assert h.adapter_op_apply(op, x.data_ptr(), y.data_ptr()) == 0
    # tet10: the reference's truncated quadrature points put it 1.5e-12 from the exact integral (tests/test_p2_tet_core.py)
    assert _rel(y.cpu().numpy(), ref) < (1e-11 if case == "tet10" else 1e-12)
    d3 = torch.zeros(nn * 3, dtype=torch.float64, device="cuda")
    assert h.adapter_op_hessian_diag(op, d3.data_ptr()) == 0 and float(d3.min()) > 0
    if case in ("hex8", "sshex8"):  # tet4 has one too (6 x, tests/test_gpu_zz_tet4_block_diag.py); tet10 / macro_tet4 have none in the reference
        b6 = torch.zeros(nn * 6, dtype=torch.float64, device="cuda")
        assert h.adapter_op_hessian_block_diag_sym(op, b6.data_ptr()) == 0
        assert torch.allclose(b6.view(nn, 6)[:, [0, 3, 5]].reshape(-1), d3, rtol=1e-13, atol=0)
    h.adapter_op_destroy(op)
    h.adapter_space_destroy(sp)


@pytest.mark.gpu
def test_derefine_op_gives_the_coarse_operator(oracles):
    """Op::derefine_op (multigrid coarse levels): sshex8 level 4 -> its macro hex8 mesh."""
    import torch

    from sfem_b200 import mesh

    h = _lib()
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    level = 4
    m = mesh.sshex8_box(level, 3, 2, 2)
    mpts = mesh.shear_points(m["m_points"])
    nn = m["n_nodes"]
    pts = np.zeros((3, nn), dtype=np.float32)
    corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
    for k, c in enumerate(corners):
        pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
    fine = _space(h, 800, level, m["elements"], pts, 1)
    coarse = _space(h, 8, 0, m["m_elements"], mpts, 1)
    op = h.adapter_create_op(fine, b"Laplacian")
    cop = h.adapter_op_derefine(op, coarse)
    assert cop and h.adapter_op_name(cop) == b"gpu:Laplacian"
    nc = mpts.shape[1]
    u = np.random.default_rng(6).uniform(-1, 1, nc)
    ref = orc.hex8_laplacian_apply_opt(m["m_elements"], orc.hex8_fff_fill(m["m_elements"], mpts), u)
    x = torch.from_numpy(u).cuda()
    y = torch.zeros_like(x)
    assert h.adapter_op_apply(cop, x.data_ptr(), y.data_ptr()) == 0 and _rel(y.cpu().numpy(), ref) < 1e-6
    for o in (cop, op):
        h.adapter_op_destroy(o)
    for s in (coarse, fine):
        h.adapter_space_destroy(s)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["hex8", "sshex8"])
def test_factory_creates_our_vector_laplacian_and_ss_lumped_mass(oracles, case):
    """sfem::create_op(space, "VectorLaplacian", DEVICE) on hex8 and sshex8 spaces with block_size 3 ("gpu:" / "ss:gpu:" keys), and
    "LumpedMass" on the sshex8 space ("ss:gpu:LumpedMass"), against the reference's own routines."""
    import torch

    from sfem_b200 import mesh

    h = _lib()
    orc = oracles.get(("reference", "f64"))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    vs = 3
    if case == "hex8":
        el, pts = mesh.box_hex8(6, 5, 7)
        pts = mesh.shear_points(pts, 0.25, 0.5, 0.125)
        nn = pts.shape[1]
        u = np.random.default_rng(8).uniform(-1, 1, nn * vs)
        ref = orc.hex8_vector_laplacian_apply(el, orc.hex8_fff_fill(el, pts), vs, u, nn)
        sp = _space(h, 8, 0, el, pts, vs)
        key = b"gpu:VectorLaplacian"
    else:
        level = 4
        m = mesh.sshex8_box(level, 2, 3, 2)
        mpts = mesh.shear_points(m["m_points"], 0.25, 0.5, 0.125)
        nn = m["n_nodes"]
        u = np.random.default_rng(9).uniform(-1, 1, nn * vs)
        ref = orc.sshex8_vector_laplacian_apply(level, m["elements"], orc.hex8_fff_fill(m["m_elements"], mpts), vs, u)
        pts = np.zeros((3, nn), dtype=np.float32)
        corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
        for k, c in enumerate(corners):
            pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
        el = m["elements"]
        sp = _space(h, 800, level, el, pts, vs)
        key = b"ss:gpu:VectorLaplacian"
    op = h.adapter_create_op(sp, b"VectorLaplacian")
    assert op and h.adapter_op_name(op) == key and h.adapter_op_n_dofs(op) == nn * vs
    x = torch.from_numpy(u).cuda()
    y = torch.zeros_like(x)
    assert h.adapter_op_apply(op, x.data_ptr(), y.data_ptr()) == 0
    assert _rel(y.cpu().numpy(), ref) < 1e-6  # fp32 FFF from b200fem_fff_fill against the reference's fill
    assert h.adapter_op_apply(op, x.data_ptr(), y.data_ptr()) == 0
    assert _rel(y.cpu().numpy(), 2 * ref) < 1e-6
    h.adapter_op_destroy(op)
    h.adapter_space_destroy(sp)
    if case == "sshex8":
        sp1 = _space(h, 800, level, el, pts, 1)
        lm = h.adapter_create_op(sp1, b"LumpedMass")
        assert lm and h.adapter_op_name(lm) == b"ss:gpu:LumpedMass"
        d = torch.zeros(nn, dtype=torch.float64, device="cuda")
        assert h.adapter_op_hessian_diag(lm, d.data_ptr()) == 0
        assert _rel(d.cpu().numpy(), orc.sshex8_lumped_mass(level, el, pts, nn)) < 1e-12
        h.adapter_op_destroy(lm)
        h.adapter_space_destroy(sp1)

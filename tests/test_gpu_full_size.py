"""GPU checks at BASELINE.json's FULL sizes through size-independent properties (the oracle needs ~0.4 s per apply at
C3 on 16 cores, so one direct comparison per configuration is affordable too):

  * constants / rigid-body motions are in the kernel of the operators,
  * symmetry  (u, A v) == (v, A u),
  * linearity A (a u + b v) == a A u + b A v,
  * the kernel variants of one plan (streaming / TMA / explicit-index; patch / atomic scatter) agree,
  * CG drives the residual down monotonically in the energy norm (a few iterations),
  * one direct oracle comparison of the full-size apply.
Bars: 1e-12 relative (fp64)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    assert torch.cuda.is_available()
    return torch


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def _properties(torch, apply_overwrite, n, tol=1e-12, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    u = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    v = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    Au, Av, Aw = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
    apply_overwrite(u, Au)
    apply_overwrite(v, Av)
    # symmetry
    uAv, vAu = float(torch.dot(u, Av)), float(torch.dot(v, Au))
    assert abs(uAv - vAu) <= 10 * tol * float(Au.norm() * v.norm()), (uAv, vAu)
    # linearity
    a, b = 0.375, -1.25
    apply_overwrite(a * u + b * v, Aw)
    assert _rel(Aw, a * Au + b * Av) < tol
    # positive semi-definite
    assert float(torch.dot(u, Au)) > 0
    return u, Au


def test_c3_sshex8_laplacian_full_size(oracles):
    """config 2 of BASELINE.json: level 8 on a 64^3 macro mesh, 135 005 697 dofs"""
    torch = _torch()
    from sfem_b200 import lib, mesh

    L, M = 8, 64
    m = mesh.sshex8_box(L, M, M, M, with_points=False)
    nn = m["n_nodes"]
    assert nn == 135_005_697
    fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, torch.float64, level=L)
    assert op.get_option("mode") == 1 and op.get_option("interior_exclusive") == 1 and op.get_option("tma_ok") == 1
    u, Au = _properties(torch, op.apply_overwrite, nn)
    # constants are in the kernel
    y = torch.empty_like(u)
    op.apply_overwrite(torch.ones_like(u), y)
    assert float(y.abs().max()) < 1e-12
    # the kernel variants of the same plan agree
    op.set_option("use_tma", 1)
    op.apply_overwrite(u, y)
    assert _rel(y, Au) < 1e-14
    op.set_option("use_tma", 0)
    # accumulate form
    y.copy_(Au)
    op.apply(u, y)
    assert _rel(y, 2 * Au) < 1e-14
    # one direct comparison with the reference's own OpenMP kernel
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    ref = orc.sshex8_laplacian_apply(L, m["elements"], fff, u.cpu().numpy())
    assert np.linalg.norm(Au.cpu().numpy() - ref) < 1e-12 * np.linalg.norm(ref)


def test_c2_hex8_linear_elasticity_full_size(oracles):
    """config 1: 128^3 hex8 elements, 6 440 067 dofs"""
    torch = _torch()
    from sfem_b200 import lib, mesh

    el, pts = mesh.box_hex8(129, 129, 129)
    pts = mesh.shear_points(pts)
    nn = pts.shape[1]
    assert nn * 3 == 6_440_067
    adj, det = lib.adjugate_and_det_fill(lib.HEX8, el, pts)
    op = lib.linear_elasticity(lib.HEX8, el, adj, det, 1.0, 1.0, nn, torch.float64)
    u, Au = _properties(torch, op.apply_overwrite, nn * 3)
    # rigid translations and an infinitesimal rotation are in the kernel
    y = torch.empty_like(u)
    p = torch.from_numpy(pts.astype(np.float64)).cuda()
    t = torch.tensor([0.3, -1.2, 0.5], dtype=torch.float64, device="cuda").repeat(nn)
    op.apply_overwrite(t, y)
    assert float(y.abs().max()) < 1e-11
    rot = torch.stack([-p[1], p[0], torch.zeros_like(p[0])], dim=1).reshape(-1).contiguous()  # omega = e_z
    op.apply_overwrite(rot, y)
    # only up to the fp32 rounding of the cached adjugates (geom_t), which breaks exact skew-symmetry of the gradient
    assert float(y.abs().max()) < 1e-7
    # atomic scatter and the other patch sizes agree with the default (staged patch kernel, 128-element patches)
    assert op.get_option("scatter") == lib.SCATTER_PATCH
    op.set_option("scatter", lib.SCATTER_ATOMIC)
    op.apply_overwrite(u, y)
    assert _rel(y, Au) < 1e-13
    for pe in (256, 512):
        op.set_option("patch_elements", pe)
        op.set_option("scatter", lib.SCATTER_PATCH)
        op.apply_overwrite(u, y)
        assert _rel(y, Au) < 1e-13
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    ref = orc.hex8_linear_elasticity_apply(el, adj, det, 1.0, 1.0, u.cpu().numpy(), nn)
    assert np.linalg.norm(Au.cpu().numpy() - ref) < 1e-12 * np.linalg.norm(ref)


def test_c1_tet4_laplacian_full_size(oracles):
    """config 0: the ~1 M-element unit cube (56^3 nodes, 998 250 tets)"""
    torch = _torch()
    from sfem_b200 import lib, mesh

    el, pts = mesh.box_tet4(56, 56, 56)
    assert el.shape[1] == 998_250 and pts.shape[1] == 175_616
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    fff = lib.fff_fill(lib.TET4, el, pts)
    op = lib.laplacian(lib.TET4, el, fff, nn, torch.float64)
    u, Au = _properties(torch, op.apply_overwrite, nn)
    y = torch.empty_like(u)
    op.apply_overwrite(torch.ones_like(u), y)
    assert float(y.abs().max()) < 1e-12
    op.set_option("scatter", lib.SCATTER_ATOMIC)
    op.apply_overwrite(u, y)
    assert _rel(y, Au) < 1e-13
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    ref = orc.tet4_laplacian_apply_opt(el, fff, u.cpu().numpy())  # same FFF array on both sides
    assert np.linalg.norm(Au.cpu().numpy() - ref) < 1e-12 * np.linalg.norm(ref)


def test_cg_full_size_energy_decrease():
    """CG on the C3 Poisson problem: the energy norm of the error decreases monotonically, i.e. the recurrence residual
    history of the host-layer CG is consistent with true residuals at full size (20 iterations)."""
    torch = _torch()
    from sfem_b200 import host, lib, mesh

    L, M = 8, 32
    m = mesh.sshex8_box(L, M, M, M, with_points=False)
    nn = m["n_nodes"]
    fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
    f = host.Function(torch.float64, nn, 1)
    f.add_operator(lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, torch.float64, level=L))
    f.add_dirichlet(mesh.sshex8_face_nodes(L, m["elements"], m["m_points"], 0, 0.0), 0, -1.0)
    f.add_dirichlet(mesh.sshex8_face_nodes(L, m["elements"], m["m_points"], 0, 1.0), 0, 1.0)
    blas = host.Blas(torch.float64)
    x = torch.zeros(nn, dtype=torch.float64, device="cuda")
    b = torch.zeros(nn, dtype=torch.float64, device="cuda")
    f.apply_constraints(x)
    f.apply_constraints(b)
    cg = host.cg(blas, f, None, rtol=1e-30, atol=1e-30, max_it=20)
    assert host.apply(cg, b, x) == 1 and host.iterations(cg) == 20  # max_it reached: SFEM_FAILURE like the reference
    h = host.residual_history(cg)
    r = torch.empty_like(x)
    host.apply_overwrite(f, x, r)
    true_res = float((b - r).norm())
    assert abs(true_res - h[-1]) < 1e-10 * h[0]
    assert h[-1] < h[0]


def test_c3_sshex8_laplacian_full_size_fp32(oracles):
    """config 2 in fp32 (BASELINE: "fp32 and fp64"): the full-size apply against the reference's fp32 build, 1e-5."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", "f32")) or oracles.get(("port", "f32"))
    if orc is None:
        pytest.skip("no fp32 oracle built")
    L, M = 8, 64
    m = mesh.sshex8_box(L, M, M, M, with_points=False)
    nn = m["n_nodes"]
    fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, torch.float32, level=L)
    g = torch.Generator(device="cuda").manual_seed(11)
    u = torch.rand(nn, dtype=torch.float32, device="cuda", generator=g) - 0.5
    Au = torch.empty_like(u)
    op.apply_overwrite(u, Au)
    ref = orc.sshex8_laplacian_apply(L, m["elements"], fff, u.cpu().numpy())
    err = np.linalg.norm(Au.cpu().numpy().astype(np.float64) - ref.astype(np.float64)) / np.linalg.norm(ref.astype(np.float64))
    print(f"C3 fp32 full size: rel-L2 vs the reference's fp32 kernel = {err:.3e}")
    assert err < 1e-5, err
    # both fp32 kernels of the plan agree, and fp32 agrees with the fp64 operator on the same input to fp32 rounding
    for use_ws in (0, 2):
        op.set_option("use_ws", use_ws)
        y = torch.empty_like(u)
        op.apply_overwrite(u, y)
        assert float((y.double() - Au.double()).norm() / Au.double().norm()) < 1e-6
    y = torch.empty_like(u)
    op.apply_overwrite(torch.ones_like(u), y)
    assert float(y.abs().max()) < 1e-4  # constants are in the kernel up to fp32 cancellation

"""GPU parity tests proper: every operator of the path, through the C ABI, against the oracle on the
same seeded inputs. Bars (BASELINE.json north_star): relative L2 <= 1e-12 in fp64, <= 1e-5 in fp32."""
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
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _fields(points, seed):
    from sfem_b200 import mesh

    nn = points.shape[1]
    return {
        "ones": np.ones(nn),
        "quadratic": mesh.quadratic_field(points),
        "random": np.random.default_rng(seed).uniform(-1, 1, nn),
    }


def _orc(oracles, real):
    return oracles.get(("reference", real)) or oracles[("port", real)]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("scatter", ["patch", "patch1024", "atomic", "colored"])
def test_tet4_laplacian(oracles, real, scatter, monkeypatch):
    torch = _torch()
    from sfem_b200 import lib, mesh

    # meshes up to 4e6 tets take 256-tet patches, larger ones 1024-tet patches: force the large-mesh plan here as well
    if scatter == "patch1024":
        monkeypatch.setenv("B200FEM_TET4_PATCH_E", "1024")
        scatter = "patch"
    else:
        monkeypatch.delenv("B200FEM_TET4_PATCH_E", raising=False)
    orc = _orc(oracles, real)
    el, pts = mesh.box_tet4(9, 7, 8)
    pts = mesh.warp_points(pts)
    fff = orc.tet4_fff_fill(el, pts)
    nn = pts.shape[1]
    op = lib.laplacian(lib.TET4, el, fff, nn, _dt(real))
    assert op.get_option("scatter") == lib.SCATTER_PATCH  # the default for tet4 (24 tets per node)
    op.set_option("scatter", {"patch": lib.SCATTER_PATCH, "colored": lib.SCATTER_COLORED, "atomic": lib.SCATTER_ATOMIC}[scatter])
    for name, u in _fields(pts, 3).items():
        ref = orc.tet4_laplacian_apply_opt(el, fff, u)
        x = torch.from_numpy(u).to(_dt(real)).cuda()
        y = torch.zeros_like(x)
        op.apply(x, y)
        if name == "ones":
            assert np.abs(y.cpu().numpy()).max() < (1e-12 if real == "f64" else 1e-4)
        else:
            assert _rel(y.cpu().numpy(), ref) < TOL[real], name
    # accumulate semantics: a second apply doubles the result
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    # overwrite form
    y.fill_(7.0)
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    # diagonal
    d = torch.zeros(nn, dtype=_dt(real), device="cuda")
    op.diag(d)
    assert _rel(d.cpu().numpy(), orc.tet4_laplacian_diag_opt(el, fff, nn)) < TOL[real]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("scatter", ["patch", "atomic", "colored"])
def test_hex8_laplacian(oracles, real, scatter):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    el, pts = mesh.box_hex8(7, 6, 5)
    pts = mesh.shear_points(pts)
    fff = orc.hex8_fff_fill(el, pts)
    nn = pts.shape[1]
    op = lib.laplacian(lib.HEX8, el, fff, nn, _dt(real))
    assert op.get_option("scatter") == lib.SCATTER_ATOMIC  # the default for hex8
    op.set_option("scatter", {"patch": lib.SCATTER_PATCH, "colored": lib.SCATTER_COLORED, "atomic": lib.SCATTER_ATOMIC}[scatter])
    for name, u in _fields(pts, 4).items():
        ref = orc.hex8_laplacian_apply_opt(el, fff, u)
        x = torch.from_numpy(u).to(_dt(real)).cuda()
        y = torch.zeros_like(x)
        op.apply(x, y)
        if name == "ones":
            assert np.abs(y.cpu().numpy()).max() < (1e-12 if real == "f64" else 1e-4)
        else:
            assert _rel(y.cpu().numpy(), ref) < TOL[real], name


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("scatter", ["patch", "atomic", "colored"])
def test_hex8_linear_elasticity(oracles, real, scatter):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    el, pts = mesh.box_hex8(6, 5, 7)
    pts = mesh.shear_points(pts)
    adj, det = orc.hex8_adjugate_and_det_fill(el, pts)
    nn = pts.shape[1]
    mu, lam = 1.3, 0.7
    op = lib.linear_elasticity(lib.HEX8, el, adj, det, mu, lam, nn, _dt(real))
    assert op.get_option("scatter") == lib.SCATTER_PATCH  # hex8 elasticity: the staged patch kernel, 128-element patches
    op.set_option("scatter", {"patch": lib.SCATTER_PATCH, "colored": lib.SCATTER_COLORED, "atomic": lib.SCATTER_ATOMIC}[scatter])
    assert op.block_size == 3 and op.rows == nn * 3
    u = np.random.default_rng(5).uniform(-1, 1, nn * 3)
    ref = orc.hex8_linear_elasticity_apply(el, adj, det, mu, lam, u, nn)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    # rigid translation is in the kernel
    t = np.tile(np.array([0.3, -1.2, 0.5]), nn)
    y.zero_()
    op.apply(torch.from_numpy(t).to(_dt(real)).cuda(), y)
    assert np.abs(y.cpu().numpy()).max() < (1e-12 if real == "f64" else 1e-4)


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 4, 8])
@pytest.mark.parametrize("path", ["ws", "ws_r1", "ws_rolled", "ws_fo1", "ws_fo2", "tma", "stream", "explicit"])
def test_sshex8_laplacian(oracles, real, level, path):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    # 23 macro elements: not a multiple of any batch size -> ragged last batch
    m = mesh.sshex8_box(level, 23, 1, 1) if level == 2 else mesh.sshex8_box(level, 3, 2, 3)
    # the warp-specialised kernels: the default (two barriers, static phases), the round-1 kernel, the two-barrier kernel with
    # the rolled z loop, and the two lateral output buffers (sfem_b200/csrc/sshex8_laplacian_ws.cu: ws_cfg)
    ws_cfg = {"ws": 11, "ws_r1": 0, "ws_rolled": 9, "ws_fo1": 16, "ws_fo2": 18}.get(path)
    if ws_cfg is not None:
        if level != 8:
            pytest.skip("the warp-specialised kernels serve level 8")
        m = mesh.sshex8_box(level, 5, 2, 2)  # 20 elements: 6 groups of 3 + a partial group of 2
    mpts = mesh.shear_points(m["m_points"])
    fff = orc.hex8_fff_fill(m["m_elements"], mpts)
    nn = m["n_nodes"]
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, _dt(real), level=level)
    assert op.get_option("mode") == 1 and op.get_option("interior_exclusive") == 1
    assert op.get_option("tma_ok") == 1  # every entity box of the generator's numbering is a contiguous id run
    op.set_option("use_ws", 2 if ws_cfg is not None else 0)  # 2 = wherever it applies (fp32 too)
    if ws_cfg is not None:
        op.set_option("ws_cfg", ws_cfg)
    op.set_option("explicit_index", int(path == "explicit"))
    op.set_option("use_tma", int(path == "tma"))  # the TMA kernel serves level >= 4; level 2 streams either way
    u = np.random.default_rng(level).uniform(-1, 1, nn)
    ref = orc.sshex8_laplacian_apply(level, m["elements"], fff, u)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    op.apply(x, y)  # accumulate
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    y.fill_(3.0)
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    # vectors that are not 16-byte aligned (views shifted by one entry) must give the same answer
    xs = torch.zeros(nn + 1, dtype=_dt(real), device="cuda")
    ys = torch.full((nn + 1,), 7.0, dtype=_dt(real), device="cuda")
    xs[1:] = x
    op.apply_overwrite(xs[1:], ys[1:])
    assert _rel(ys[1:].cpu().numpy(), ref) < TOL[real] and float(ys[0]) == 7.0
    ones = torch.ones_like(x)
    y.zero_()
    op.apply(ones, y)
    assert np.abs(y.cpu().numpy()).max() < (1e-12 if real == "f64" else 1e-4)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_sshex8_laplacian_element_ranges(oracles, real):
    """b200fem_op_apply_range: any split of [0, ne) into ranges gives the whole-mesh result, in both the accumulate
    and the overwrite protocol (mode 1 then 2) -- what the overlapped distributed apply relies on."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    L = 8
    m = mesh.sshex8_box(L, 3, 2, 3)
    fff = orc.hex8_fff_fill(m["m_elements"], mesh.shear_points(m["m_points"]))
    nn, ne = m["n_nodes"], m["elements"].shape[1]
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, _dt(real), level=L)
    u = np.random.default_rng(3).uniform(-1, 1, nn)
    ref = orc.sshex8_laplacian_apply(L, m["elements"], fff, u)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.full((nn,), 5.0, dtype=_dt(real), device="cuda")
    op.set_option("reserve_sms", 8)
    op.apply_range(x, y, 6, 13, 1)   # zeroes the shared rows, middle range first
    op.set_option("reserve_sms", 0)
    op.apply_range(x, y, 0, 6, 2)
    op.apply_range(x, y, 13, ne, 2)
    op.apply_range(x, y, 4, 4, 2)    # empty range
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    for a, b in ((0, 5), (5, 6), (6, ne)):
        op.apply_range(x, y, a, b, 0)
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    with pytest.raises(lib.B200FemError):
        op.apply_range(x, y, 3, ne + 1, 0)
    tet_el, tet_pts = mesh.box_tet4(3, 3, 3)
    tet = lib.laplacian(lib.TET4, tet_el, orc.tet4_fff_fill(tet_el, tet_pts), tet_pts.shape[1], _dt(real))
    with pytest.raises(lib.B200FemError, match="not implemented"):
        tet.apply_range(x[: tet_pts.shape[1]], y[: tet_pts.shape[1]], 0, 1, 0)


def test_sshex8_laplacian_in_kernel_ghost_wait(oracles):
    """b200fem_op_apply_overwrite_after: the element groups from e_dependent on wait inside the kernel for a device flag (the
    distributed apply hands it the halo's ready flag). With the flag already at its value, and with the flag set by a second
    stream while the kernel is running, the result is the plain y = A x."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, "f64")
    L = 8
    m = mesh.sshex8_box(L, 5, 4, 3)
    fff = orc.hex8_fff_fill(m["m_elements"], mesh.shear_points(m["m_points"]))
    nn, ne = m["n_nodes"], m["elements"].shape[1]
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, torch.float64, level=L)
    assert op.get_option("ghost_wait") == 1
    u = np.random.default_rng(5).uniform(-1, 1, nn)
    ref = orc.sshex8_laplacian_apply(L, m["elements"], fff, u)
    x = torch.from_numpy(u).cuda()
    flag = torch.full((1,), 7, dtype=torch.int64, device="cuda")
    y = torch.full((nn,), 3.0, dtype=torch.float64, device="cuda")
    op.apply_overwrite_after(x, y, ne // 2, flag, 7)          # already there
    assert _rel(y.cpu().numpy(), ref) < TOL["f64"]
    op.apply_overwrite_after(x, y, 0, None, 0)               # no flag: the plain apply
    assert _rel(y.cpu().numpy(), ref) < TOL["f64"]
    # the flag arrives from another stream after the kernel has been launched
    y.fill_(3.0)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):  # load the two kernels now: a first launch (lazy module loading) may wait for running kernels
        torch.cuda._sleep(1000)
        flag.fill_(7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    op.apply_overwrite_after(x, y, 3, flag, 8)
    e1.record()
    with torch.cuda.stream(side):
        torch.cuda._sleep(2_000_000)  # ~1 ms on the SM the apply leaves free
        flag.fill_(8)
    torch.cuda.synchronize()
    assert _rel(y.cpu().numpy(), ref) < TOL["f64"]
    ms = e0.elapsed_time(e1)
    assert 0.3 < ms < 1000.0, f"{ms} ms: the kernel neither waited for the flag nor ran into its 2 s time-out"
    tet_el, tet_pts = mesh.box_tet4(3, 3, 3)
    tet = lib.laplacian(lib.TET4, tet_el, orc.tet4_fff_fill(tet_el, tet_pts), tet_pts.shape[1], torch.float64)
    with pytest.raises(lib.B200FemError):
        tet.apply_overwrite_after(x[: tet_pts.shape[1]], y[: tet_pts.shape[1]], 0, flag, 9)


def test_sshex8_laplacian_hierarchical_numbering(oracles):
    """Non-affine numbering (coarse lattices first) must take the explicit-index path and still match."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, "f64")
    L = 4
    m = mesh.sshex8_box(L, 2, 3, 2)
    el, mapping, _ = mesh.sshex8_hierarchical_renumbering(L, m["elements"], m["n_nodes"])
    fff = orc.hex8_fff_fill(m["m_elements"], mesh.shear_points(m["m_points"]))
    nn = m["n_nodes"]
    op = lib.laplacian(lib.SSHEX8, el, fff, nn, torch.float64, level=L)
    assert op.get_option("mode") == 0
    u = np.random.default_rng(9).uniform(-1, 1, nn)
    ref = orc.sshex8_laplacian_apply(L, el, fff, u)
    x = torch.from_numpy(u).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < 1e-12


def test_sshex8_laplacian_diag_matches_unit_applies(oracles):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, "f64")
    L = 4
    m = mesh.sshex8_box(L, 2, 1, 1)
    fff = orc.hex8_fff_fill(m["m_elements"], mesh.shear_points(m["m_points"]))
    nn = m["n_nodes"]
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, torch.float64, level=L)
    d = torch.zeros(nn, dtype=torch.float64, device="cuda")
    op.diag(d)
    ref = np.zeros(nn)
    for i in range(nn):
        e = np.zeros(nn)
        e[i] = 1
        ref[i] = orc.sshex8_laplacian_apply(L, m["elements"], fff, e)[i]
    assert _rel(d.cpu().numpy(), ref) < 1e-12


def test_geometry_fill(oracles):
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, "f64")
    el, pts = mesh.box_tet4(6, 5, 4)
    pts = mesh.warp_points(pts)
    assert _rel(lib.fff_fill(lib.TET4, el, pts), orc.tet4_fff_fill(el, pts)) < 1e-5
    el, pts = mesh.box_hex8(6, 5, 4)
    pts = mesh.warp_points(pts)
    assert _rel(lib.fff_fill(lib.HEX8, el, pts), orc.hex8_fff_fill(el, pts)) < 1e-6
    adj, det = lib.adjugate_and_det_fill(lib.HEX8, el, pts)
    radj, rdet = orc.hex8_adjugate_and_det_fill(el, pts)
    assert _rel(adj, radj) < 1e-6 and _rel(det, rdet) < 1e-6


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("n", [0, 1, 7, 1000, 1 << 20 | 5])
def test_blas(real, n):
    torch = _torch()
    from sfem_b200.lib import blas

    dt = _dt(real)
    tol = 1e-14 if real == "f64" else 1e-6
    g = torch.Generator(device="cuda").manual_seed(n)
    x = torch.rand(n, dtype=dt, device="cuda", generator=g) - 0.5
    y = torch.rand(n, dtype=dt, device="cuda", generator=g) - 0.5
    z = torch.rand(n, dtype=dt, device="cuda", generator=g) - 0.5
    xn, yn, zn = (t.cpu().numpy().astype(np.float64) for t in (x, y, z))

    assert abs(blas.dot(x, y) - float(xn @ yn)) <= tol * max(1.0, float(np.abs(xn) @ np.abs(yn)))
    assert abs(blas.norm2(x) - float(np.sqrt(xn @ xn))) <= tol * max(1.0, float(np.sqrt(xn @ xn)))

    def close(t, ref):
        return n == 0 or _rel(t.cpu().numpy(), ref) < tol

    t = y.clone(); blas.axpy(0.7, x, t); assert close(t, yn + 0.7 * xn)
    t = y.clone(); blas.axpby(0.7, x, -1.1, t); assert close(t, 0.7 * xn - 1.1 * yn)
    t = z.clone(); blas.zaxpby(0.7, x, -1.1, y, t); assert close(t, 0.7 * xn - 1.1 * yn)
    t = z.clone(); blas.xypaz(x, y, 0.3, t); assert close(t, xn * yn + 0.3 * zn)
    t = x.clone(); blas.scal(-2.5, t); assert close(t, -2.5 * xn)
    t = x.clone(); blas.values(1.25, t); assert close(t, np.full(n, 1.25))
    t = x.clone(); blas.zeros(t); assert n == 0 or float(t.abs().max()) == 0.0
    t = torch.empty_like(x); blas.copy(x, t); assert n == 0 or torch.equal(t, x)
    t = x.clone()
    if n > 2:
        t[1] = 0.0
    blas.reciprocal(2.0, t)
    ref = np.where(xn != 0, 2.0 / np.where(xn != 0, xn, 1), 0.0)
    if n > 2:
        ref[1] = 0.0
    assert close(t, ref)
    # unaligned views take the scalar path
    if n > 8:
        xv, yv = x[1:], y[1:]
        assert abs(blas.dot(xv, yv) - float(xn[1:] @ yn[1:])) <= tol * max(1.0, float(np.abs(xn) @ np.abs(yn)))
        t = yv.clone(); blas.axpy(0.5, xv, t); assert close(t, yn[1:] + 0.5 * xn[1:])

    # fused CG / Chebyshev updates
    p, Ap, xx, r = x.clone(), y.clone(), z.clone(), (x * y).clone()
    pn, An, xxn, rn = xn, yn, zn, (x * y).cpu().numpy().astype(np.float64)
    rtr = blas.cg_update_x_r_rtr(0.37, p, Ap, xx, r)
    rr = rn - 0.37 * An
    assert close(xx, xxn + 0.37 * pn) and close(r, rr)
    assert abs(rtr - float(rr @ rr)) <= tol * max(1.0, float(rr @ rr))
    t = x.clone(); blas.cg_update_p(0.9, y, t); assert close(t, yn + 0.9 * xn)
    pp, xs = x.clone(), z.clone()
    blas.cheb3_update(0.2, 0.6, y, (x * y), pp, xs)
    pr = 0.6 * xn - yn + rn
    assert close(pp, pr) and close(xs, zn - 0.2 * pr)


def test_dot_is_reproducible():
    torch = _torch()
    from sfem_b200.lib import blas

    x = torch.rand(3_000_001, dtype=torch.float64, device="cuda")
    y = torch.rand(3_000_001, dtype=torch.float64, device="cuda")
    vals = {blas.dot(x, y) for _ in range(5)}
    assert len(vals) == 1


def test_dirichlet_kernels():
    torch = _torch()
    from sfem_b200 import lib

    n, bs = 1000, 3
    src = torch.rand(n * bs, dtype=torch.float64, device="cuda")
    dst = torch.zeros_like(src)
    idx = torch.tensor([0, 5, 17, 999], dtype=torch.int32, device="cuda")
    lib.constraint_nodes_copy_vec(idx, bs, 1, src, dst)
    ref = np.zeros(n * bs)
    ii = idx.cpu().numpy() * bs + 1
    ref[ii] = src.cpu().numpy()[ii]
    assert np.array_equal(dst.cpu().numpy(), ref)
    lib.constraint_nodes_to_value_vec(idx, bs, 2, -4.5, dst)
    ref[idx.cpu().numpy() * bs + 2] = -4.5
    assert np.array_equal(dst.cpu().numpy(), ref)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_golden_vectors_on_gpu(real):
    """The CUDA path against the committed reference outputs (tests/golden, generated from oracle/_ref): needs no
    oracle library on the box."""
    import os

    torch = _torch()
    from sfem_b200 import lib

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    dt = _dt(real)

    def run(op, u):
        x = torch.from_numpy(np.asarray(u)).to(dt).cuda()
        y = torch.zeros_like(x)
        op.apply(x, y)
        return y.cpu().numpy()

    g = np.load(os.path.join(gold, "tet4_laplacian.npz"))
    nn = g["points"].shape[1]
    assert _rel(run(lib.laplacian(lib.TET4, g["elements"], g["fff"], nn, dt), g["u"]), g[f"y_{real}"]) < TOL[real]
    g = np.load(os.path.join(gold, "hex8.npz"))
    nn = g["points"].shape[1]
    assert _rel(run(lib.laplacian(lib.HEX8, g["elements"], g["fff"], nn, dt), g["u"]), g[f"y_lap_{real}"]) < TOL[real]
    op = lib.linear_elasticity(lib.HEX8, g["elements"], g["adjugate"], g["determinant"], float(g["mu"]), float(g["lam"]), nn, dt)
    assert _rel(run(op, g["u3"]), g[f"y_le_{real}"]) < TOL[real]
    for L in (2, 4, 8):
        g = np.load(os.path.join(gold, f"sshex8_L{L}.npz"))
        nn = g["u"].size
        op = lib.laplacian(lib.SSHEX8, g["elements"], g["fff"], nn, dt, level=L)
        assert _rel(run(op, g["u"]), g[f"y_lap_{real}"]) < TOL[real], L


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("family", ["tet10", "macro_tet4"])
def test_golden_p2_tet_on_gpu(real, family):
    """10-node tetrahedra against the committed reference outputs (tests/golden/p2_tet.npz, `make_golden.py p2_tet`): Laplacian and
    linear elasticity, apply + diag, both scatters, the packed Laplacian, the tet10 lumped mass -- no oracle library needed."""
    import os

    torch = _torch()
    from sfem_b200 import lib

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "p2_tet.npz"))
    dt = _dt(real)
    el, nn = g["elements"], g["points"].shape[1]
    et = lib.TET10 if family == "tet10" else lib.MACRO_TET4
    le_tol = (1e-11 if family == "tet10" else 1e-12) if real == "f64" else 1e-5  # tet10: see test_p2_tet_linear_elasticity

    def run(op, u):
        x = torch.from_numpy(np.asarray(u)).to(dt).cuda()
        y = torch.zeros_like(x)
        op.apply(x, y)
        return y.cpu().numpy()

    lap = lib.laplacian(et, el, g["fff"], nn, dt)
    le = lib.linear_elasticity(et, el, g["adjugate"], g["determinant"], float(g["mu"]), float(g["lam"]), nn, dt)
    for sc in (lib.SCATTER_ATOMIC, lib.SCATTER_PATCH):
        lap.set_option("scatter", sc)
        le.set_option("scatter", sc)
        assert _rel(run(lap, g["u"]), g[f"{family}_lap_{real}"]) < TOL[real], sc
        assert _rel(run(le, g["u3"]), g[f"{family}_le_{real}"]) < le_tol, sc
    if real == "f64":
        d = torch.zeros(nn, dtype=dt, device="cuda")
        lap.diag(d)
        assert _rel(d.cpu().numpy(), g[f"{family}_lap_diag_f64"]) < 1e-12
        d3 = torch.zeros(nn * 3, dtype=dt, device="cuda")
        le.diag(d3)
        assert _rel(d3.cpu().numpy(), g[f"{family}_le_diag_f64"]) < le_tol
        if family == "tet10":
            m = torch.zeros(nn, dtype=dt, device="cuda")
            lib.mass(lib.TET10, el, g["points"], nn, dt).diag(m)
            assert _rel(m.cpu().numpy(), g["lumped_tet10_f64"]) < 1e-10
    # packed: the same operator in the packed numbering, y_packed[i] = y[new_to_old[i]]
    pm = lib.PackedMesh(el, nn, 64)
    pop = lib.packed_laplacian(et, pm, g["fff"], dt)
    for kernel in (1, 0):
        pop.set_option("packed_kernel", kernel)
        assert _rel(run(pop, g["u"][pm.new_to_old]), g[f"{family}_lap_{real}"][pm.new_to_old]) < TOL[real], kernel


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_hex8_diagonals(oracles, real):
    """hex8 Laplacian diag, hex8 linear-elasticity diag and symmetric block diagonal against the reference's
    hex8_laplacian_diag / affine_hex8_linear_elasticity_diag / _block_diag_sym (operators/hex8/hex8_laplacian.hpp:36,
    hex8_linear_elasticity.cpp:1043-1200). Those take the points; the geometry here is dyadic (unit spacing, dyadic shear),
    so the fp32 FFF / adjugates of the plan are exactly what the reference derives."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("needs the reference build for the diagonal oracles")
    el, pts = mesh.box_hex8(6, 5, 7, 5.0, 4.0, 6.0)
    pts = mesh.shear_points(pts, 0.25, 0.5, 0.125)
    nn = pts.shape[1]
    dt = _dt(real)
    fff = orc.hex8_fff_fill(el, pts)
    op = lib.laplacian(lib.HEX8, el, fff, nn, dt)
    d = torch.zeros(nn, dtype=dt, device="cuda")
    op.diag(d)
    assert _rel(d.cpu().numpy(), orc.hex8_laplacian_diag(el, pts, nn)) < TOL[real]
    op.diag(d)  # diag accumulates like Op::hessian_diag
    assert _rel(d.cpu().numpy(), 2 * orc.hex8_laplacian_diag(el, pts, nn)) < TOL[real]
    mu, lam = 1.3, 0.7
    adj, det = orc.hex8_adjugate_and_det_fill(el, pts)
    le = lib.linear_elasticity(lib.HEX8, el, adj, det, mu, lam, nn, dt)
    d3 = torch.zeros(nn * 3, dtype=dt, device="cuda")
    le.diag(d3)
    assert _rel(d3.cpu().numpy(), orc.hex8_linear_elasticity_diag(el, pts, mu, lam, nn)) < TOL[real]
    b6 = torch.zeros(nn * 6, dtype=dt, device="cuda")
    le.block_diag_sym(b6)
    # the reference's generated block-diagonal code is itself off by 1.4e-11 (relative, uniformly): on this dyadic geometry the
    # closed form is exact (6.00703125 at node 52, the reference returns 6.007031249918575)
    assert _rel(b6.cpu().numpy(), orc.hex8_linear_elasticity_block_diag_sym(el, pts, mu, lam, nn)) < (1e-10 if real == "f64" else TOL[real])
    # the diagonal is what unit applies see
    if real == "f64":
        for i in (0, nn * 3 // 2 + 1, nn * 3 - 1):
            e = torch.zeros(nn * 3, dtype=dt, device="cuda")
            e[i] = 1.0
            y = torch.zeros_like(e)
            le.apply(e, y)
            assert abs(float(y[i]) - float(d3[i])) < 1e-12 * abs(float(d3[i]))

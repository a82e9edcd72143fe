"""GPU parity for the multigrid-side rows of SURVEY section 8(a): sshex8 linear elasticity (apply, diagonal,
block diagonal), the sshex8 transfer operators and the block-Jacobi kernels, through the C ABI, against the
oracle on the same seeded inputs. Bars: relative L2 <= 1e-12 (fp64), <= 1e-5 (fp32)."""
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


def _orc(oracles, real):
    return oracles.get(("reference", real)) or oracles[("port", real)]


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 4, 8])
@pytest.mark.parametrize("explicit", [0, 1])
def test_sshex8_linear_elasticity(oracles, real, level, explicit):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, real)
    m = mesh.sshex8_box(level, 3, 2, 3 if level < 8 else 2)
    mpts = mesh.shear_points(m["m_points"])
    adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
    nn = m["n_nodes"]
    mu, lam = 1.3, 0.7
    op = lib.linear_elasticity(lib.SSHEX8, m["elements"], adj, det, mu, lam, nn, _dt(real), level=level)
    assert op.block_size == 3 and op.rows == nn * 3
    assert op.get_option("mode") == 1  # the generator's numbering is affine per entity
    op.set_option("explicit_index", explicit)
    u = np.random.default_rng(level).uniform(-1, 1, nn * 3)
    ref = orc.sshex8_linear_elasticity_apply(level, m["elements"], adj, det, mu, lam, u, nn)
    x = torch.from_numpy(u).to(_dt(real)).cuda()
    y = torch.zeros_like(x)
    op.apply(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    op.apply(x, y)  # accumulate
    assert _rel(y.cpu().numpy(), 2 * ref) < TOL[real]
    y.fill_(3.0)
    op.apply_overwrite(x, y)
    assert _rel(y.cpu().numpy(), ref) < TOL[real]
    # rigid translations are in the kernel
    t = torch.ones(nn * 3, dtype=_dt(real), device="cuda")
    y.zero_()
    op.apply(t, y)
    assert np.abs(y.cpu().numpy()).max() < (1e-11 if real == "f64" else 1e-3)


def test_sshex8_linear_elasticity_ragged_and_empty(oracles):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = _orc(oracles, "f64")
    level = 8
    m = mesh.sshex8_box(level, 3, 1, 3)  # 9 macro elements: not a multiple of the 4-element batch
    mpts = mesh.shear_points(m["m_points"])
    adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
    nn = m["n_nodes"]
    op = lib.linear_elasticity(lib.SSHEX8, m["elements"], adj, det, 1.0, 1.0, nn, torch.float64, level=level)
    u = np.random.default_rng(0).uniform(-1, 1, nn * 3)
    ref = orc.sshex8_linear_elasticity_apply(level, m["elements"], adj, det, 1.0, 1.0, u, nn)
    y = torch.zeros(nn * 3, dtype=torch.float64, device="cuda")
    op.apply(torch.from_numpy(u).cuda(), y)
    assert _rel(y.cpu().numpy(), ref) < 1e-12
    empty = lib.linear_elasticity(lib.SSHEX8, m["elements"][:, :0], adj[:0], det[:0], 1.0, 1.0, nn, torch.float64, level=level)
    y.fill_(5.0)
    empty.apply_overwrite(torch.from_numpy(u).cuda(), y)
    assert float(y.abs().max()) == 0.0


@pytest.mark.parametrize("level", [2, 4, 8])
def test_sshex8_linear_elasticity_diagonals(oracles, level):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", "f64"))
    if orc is None:
        pytest.skip("needs the reference build for the diagonal oracles")
    m = mesh.sshex8_box(level, 2, 3, 2, 2.0, 3.0, 2.0)  # unit macro spacing
    mpts = mesh.shear_points(m["m_points"], 0.25, 0.5, 0.125)  # dyadic: fp32 geometry is exact
    adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
    nn = m["n_nodes"]
    mu, lam = 1.3, 0.7
    pts = np.zeros((3, nn), dtype=np.float32)
    corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
    for k, c in enumerate(corners):
        pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
    op = lib.linear_elasticity(lib.SSHEX8, m["elements"], adj, det, mu, lam, nn, torch.float64, level=level)
    d = torch.zeros(nn * 3, dtype=torch.float64, device="cuda")
    op.diag(d)
    # dyadic geometry: the fp32 adjugates we are handed and the ones the reference derives from the points are identical
    e_d = _rel(d.cpu().numpy(), orc.sshex8_linear_elasticity_diag(level, m["elements"], pts, mu, lam, nn))
    print(f"sshex8 LE diag L{level}: rel-L2 vs oracle = {e_d:.3e}")
    assert e_d < 1e-12, e_d
    b = torch.zeros(nn * 6, dtype=torch.float64, device="cuda")
    op.block_diag_sym(b)
    e_b = _rel(b.cpu().numpy(), orc.sshex8_linear_elasticity_block_diag_sym(level, m["elements"], pts, mu, lam, nn))
    print(f"sshex8 LE block diag L{level}: rel-L2 vs oracle = {e_b:.3e}")
    # the reference's generated block-diagonal routine carries a uniform 1.36e-11 relative error (the same one its hex8 twin
    # shows: tests/test_gpu_operators.py::test_hex8_diagonals; its plain diagonal agrees to 1e-16 above)
    assert e_b < 1e-10, e_b
    assert _rel(b.view(nn, 6)[:, [0, 3, 5]].reshape(-1).cpu().numpy(), d.cpu().numpy()) < 1e-14
    # the diagonal is what unit applies see
    for i in (0, nn * 3 // 2, nn * 3 - 1):
        e = torch.zeros(nn * 3, dtype=torch.float64, device="cuda")
        e[i] = 1.0
        y = torch.zeros_like(e)
        op.apply(e, y)
        assert abs(float(y[i]) - float(d[i])) < 1e-12 * abs(float(d[i]))


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("fine,coarse", [(8, 1), (4, 1), (2, 1), (8, 4), (8, 2), (4, 2)])
@pytest.mark.parametrize("vec_size", [1, 3])
def test_sshex8_transfer(oracles, real, fine, coarse, vec_size):
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("needs the reference build for the transfer oracles")
    m = mesh.sshex8_box(fine, 3, 2, 2)
    el, nn = m["elements"], m["n_nodes"]
    # the coarse lattice lives on the same node ids (strided view of the fine connectivity)
    stride = fine // coarse
    tr = lib.Transfer(fine, 1, el, nn, coarse, stride, el, nn)
    count = orc.sshex8_element_node_incidence_count(fine, 1, el, nn)
    assert np.array_equal(tr.incidence_count(), count)
    rng = np.random.default_rng(fine * 10 + coarse)
    fvec = rng.uniform(-1, 1, nn * vec_size)
    if coarse == 1:
        ref_r = orc.sshex8_hierarchical_restriction(fine, el, count, vec_size, fvec, nn)
    else:
        ref_r = orc.sshex8_restrict(fine, 1, el, count, coarse, stride, el, vec_size, fvec, nn)
    to = torch.zeros(nn * vec_size, dtype=_dt(real), device="cuda")
    tr.restrict(torch.from_numpy(fvec).to(_dt(real)).cuda(), to, vec_size)
    assert _rel(to.cpu().numpy(), ref_r) < TOL[real]

    cvec = rng.uniform(-1, 1, nn * vec_size)
    if coarse == 1:
        ref_p = orc.sshex8_hierarchical_prolongation(fine, el, vec_size, cvec, nn)
    else:
        ref_p = orc.sshex8_prolongate(coarse, stride, el, fine, 1, el, vec_size, cvec, nn)
    to = torch.zeros(nn * vec_size, dtype=_dt(real), device="cuda")
    tr.prolongate(torch.from_numpy(cvec).to(_dt(real)).cuda(), to, vec_size)
    assert _rel(to.cpu().numpy(), ref_p) < TOL[real]


def test_sshex8_transfer_separate_numbering(oracles):
    """coarse mesh with its own numbering (the hex8 macro mesh): the 2-level mgsolve configuration."""
    torch = _torch()
    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", "f64"))
    if orc is None:
        pytest.skip("needs the reference build")
    L = 4
    m = mesh.sshex8_box(L, 2, 2, 3)
    el, nn = m["elements"], m["n_nodes"]
    mel = m["m_elements"]
    n_macro = int(mel.max()) + 1
    # hex8 (VTK corner order) -> level-1 lattice rows
    lat = np.zeros((8, mel.shape[1]), dtype=np.int32)
    for k, (cx, cy, cz) in enumerate(mesh.HEX8_CORNERS):
        lat[mesh.sshex8_lidx(1, cx, cy, cz)] = mel[k]
    tr = lib.Transfer(L, 1, el, nn, 1, 1, lat, n_macro)
    rng = np.random.default_rng(1)
    cvec = rng.uniform(-1, 1, n_macro)
    to = torch.zeros(nn, dtype=torch.float64, device="cuda")
    tr.prolongate(torch.from_numpy(cvec).cuda(), to)
    # macro node ids are the first ids of the sshex8 numbering: compare with the hierarchical oracle
    ref = orc.sshex8_hierarchical_prolongation(L, el, 1, np.concatenate([cvec, np.zeros(nn - n_macro)]), nn)
    assert _rel(to.cpu().numpy(), ref) < 1e-12
    # R = P^T up to the incidence weights: (R (c .* f), g) == (f, P g)
    f = rng.uniform(-1, 1, nn)
    g = rng.uniform(-1, 1, n_macro)
    count = tr.incidence_count().astype(np.float64)
    r = torch.zeros(n_macro, dtype=torch.float64, device="cuda")
    tr.restrict(torch.from_numpy(f * count).cuda(), r)
    pg = torch.zeros(nn, dtype=torch.float64, device="cuda")
    tr.prolongate(torch.from_numpy(g).cuda(), pg)
    # every (element, node) pair contributes once on the restriction side: weight by incidence on the right
    lhs = float(np.dot(r.cpu().numpy(), g))
    rhs = float(np.dot(f * count, pg.cpu().numpy()))
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0)


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_block_sym_jacobi(real):
    torch = _torch()
    from sfem_b200 import lib

    rng = np.random.default_rng(3)
    nb = 1000
    a = rng.uniform(-1, 1, (nb, 3, 3))
    spd = a @ a.transpose(0, 2, 1) + 3 * np.eye(3)
    sym6 = np.stack([spd[:, 0, 0], spd[:, 0, 1], spd[:, 0, 2], spd[:, 1, 1], spd[:, 1, 2], spd[:, 2, 2]], axis=1)
    mask_bits = rng.uniform(size=nb * 3) < 0.1
    mask = np.packbits(mask_bits, bitorder="little")
    relax = 1.0 / 3.0
    # reference semantics (algebra/sfem_ShiftableJacobi.hpp:86-100)
    full = spd.copy()
    for b in range(nb):
        for r in range(3):
            if mask_bits[b * 3 + r]:
                full[b, r, :] = 0
                full[b, r, r] = 1
    inv = relax * np.linalg.inv(full)
    d6 = torch.from_numpy(sym6.reshape(-1)).to(_dt(real)).cuda()
    inv9 = torch.zeros(nb * 9, dtype=_dt(real), device="cuda")
    lib.block_sym_jacobi_setup(d6, relax, inv9, mask=torch.from_numpy(mask).cuda())
    assert _rel(inv9.cpu().numpy(), inv.reshape(-1)) < (1e-12 if real == "f64" else 1e-4)
    x = rng.uniform(-1, 1, nb * 3)
    y0 = rng.uniform(-1, 1, nb * 3)
    y = torch.from_numpy(y0).to(_dt(real)).cuda()
    lib.block_jacobi_apply(inv9, torch.from_numpy(x).to(_dt(real)).cuda(), y)
    ref = y0 + np.einsum("bij,bj->bi", inv, x.reshape(nb, 3)).reshape(-1)
    assert _rel(y.cpu().numpy(), ref) < (1e-12 if real == "f64" else 1e-4)


def test_mixed_precision_block_sym_jacobi(oracles):
    """MixedPrecisionShiftableBlockSymJacobi<double, float> (algebra/sfem_MixedPrecisionShiftableBlockSymJacobi.hpp): fp64
    vectors, fp32 inverted blocks. Against the reference's own class (oracle/_ref solver build) when it is on this box,
    and against the fp64 smoother to fp32 accuracy."""
    torch = _torch()
    from sfem_b200 import host

    rng = np.random.default_rng(4)
    nb = 2000
    a = rng.uniform(-1, 1, (nb, 3, 3))
    spd = a @ a.transpose(0, 2, 1) + 3 * np.eye(3)
    sym6 = np.stack([spd[:, 0, 0], spd[:, 0, 1], spd[:, 0, 2], spd[:, 1, 1], spd[:, 1, 2], spd[:, 2, 2]], axis=1).reshape(-1)
    mask_bits = rng.uniform(size=nb * 3) < 0.1
    mask = np.packbits(mask_bits, bitorder="little")
    relax = 1.0 / 3.0
    x = rng.uniform(-1, 1, nb * 3)
    y0 = rng.uniform(-1, 1, nb * 3)
    d6 = torch.from_numpy(sym6).cuda()
    mk = torch.from_numpy(mask).cuda()
    j = host.mixed_block_sym_jacobi(d6, mk, relax)
    y = torch.from_numpy(y0).cuda()
    assert host.apply(j, torch.from_numpy(x).cuda(), y) == 0
    # fp64 smoother of the same blocks: equal to fp32 accuracy, but NOT to fp64 accuracy (the blocks really are fp32)
    j64 = host.block_sym_jacobi(host.Blas(torch.float64), d6, mk, relax)
    y64 = torch.from_numpy(y0).cuda()
    host.apply(j64, torch.from_numpy(x).cuda(), y64)
    diff = float((y - y64).norm() / (y64 - torch.from_numpy(y0).cuda()).norm())
    assert 1e-10 < diff < 1e-5, diff
    from oracle import ref_solvers

    if ref_solvers.available("f64"):
        rs = ref_solvers.RefSolvers("f64")
        jr = rs.mixed_block_sym_jacobi(sym6, mask, relax)
        yr = y0.copy()
        assert rs.apply(jr, np.ascontiguousarray(x), yr) == 0
        # same fp32 operations on both sides; the only freedom is fp32 contraction / evaluation order inside the 3x3 inverse
        assert np.linalg.norm(y.cpu().numpy() - yr) < 2e-6 * np.linalg.norm(yr - y0)

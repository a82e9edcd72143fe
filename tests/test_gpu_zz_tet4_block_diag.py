"""tet4 linear elasticity: the symmetric 3 x 3 node blocks (b200fem_op_block_diag_sym on the TET4 plans, cached adjugate and from the
points) against tet4_linear_elasticity_block_diag_sym (operators/tet4/tet4_linear_elasticity.cpp:360-440). The reference's blocks are
6 x the blocks of its own apply (tests/test_p2_tet_core.py pins that); the plans reproduce it by default, option
"reference_block_diag" = 0 gives the blocks consistent with apply / diag. The assembly loop is also checked on the CPU
(tests/test_p2_tet_core.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("real", ["f64", "f32"])
def test_tet4_linear_elasticity_block_diag_sym(oracles, real):
    import torch

    from sfem_b200 import lib, mesh

    orc = oracles.get(("reference", real))
    if orc is None:
        pytest.skip("oracle/_ref not built")
    dt = torch.float64 if real == "f64" else torch.float32
    el, pts = mesh.box_tet4(5, 6, 4, 4.0, 5.0, 3.0)
    pts = mesh.shear_points(pts, 0.25, 0.5, 0.125)  # dyadic: the fp32 adjugate of the cached plan is exact
    nn = pts.shape[1]
    mu, lam = 1.3, 0.7
    ref = orc.tet4_linear_elasticity_block_diag_sym(el, pts, mu, lam, nn)
    tol = 1e-12 if real == "f64" else 1e-5
    rel = lambda a, b: np.linalg.norm(np.asarray(a, dtype=np.float64) - b) / np.linalg.norm(b)
    adj, det = lib.adjugate_and_det_fill(lib.TET4, el, pts)
    for op in (lib.linear_elasticity(lib.TET4, el, adj, det, mu, lam, nn, dt), lib.linear_elasticity_from_points(lib.TET4, el, pts, mu, lam, nn, dt)):
        assert op.get_option("reference_block_diag") == 1
        b6 = torch.zeros(nn * 6, dtype=dt, device="cuda")
        op.block_diag_sym(b6)
        assert rel(b6.cpu().numpy(), ref) < tol
        d3 = torch.zeros(nn * 3, dtype=dt, device="cuda")
        op.diag(d3)
        op.set_option("reference_block_diag", 0)
        b6.zero_()
        op.block_diag_sym(b6)
        assert rel(b6.cpu().numpy(), ref / 6) < tol
        assert torch.allclose(b6.view(nn, 6)[:, [0, 3, 5]].reshape(-1), d3, rtol=1e-12 if real == "f64" else 1e-5, atol=0)

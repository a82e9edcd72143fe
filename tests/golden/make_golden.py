#!/usr/bin/env python
"""Generates tests/golden/*.npz from the reference-compiled oracle (oracle/_ref, built by `make -C oracle ref`
from the sources under /root/reference). Run in the build container only; the vectors travel, the reference
does not. Every file holds the inputs and the reference's outputs of one operator on a small seeded case."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import Oracle  # noqa: E402
from sfem_b200 import mesh  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def p2_tet():
    """10-node tetrahedra (tet10 / macro_tet4): own seed, own file, so the older fixtures stay byte-identical."""
    o64, o32 = Oracle("reference", "f64"), Oracle("reference", "f32")
    rng = np.random.default_rng(20261018)
    el4, pts4 = mesh.box_tet4(4, 5, 4)
    el, pts = mesh.tet4_to_tet10(el4, mesh.warp_points(pts4))
    nn = pts.shape[1]
    fff = o64.tet4_fff_fill(el[:4], pts)
    p = pts.astype(np.float64)
    J = np.moveaxis(np.stack([p[:, el[c + 1]] - p[:, el[0]] for c in range(3)], axis=1), 2, 0)
    det = np.linalg.det(J)
    adj = (np.linalg.inv(J) * det[:, None, None]).reshape(-1).astype(np.float32)
    det = det.astype(np.float32)
    u, u3 = rng.uniform(-1, 1, nn), rng.uniform(-1, 1, nn * 3)
    d = dict(elements=el, points=pts, fff=fff, adjugate=adj, determinant=det, u=u, u3=u3, mu=1.3, lam=0.7,
             lumped_tet10_f64=o64.tet10_lumped_mass(el, pts, nn))
    for fam in ("tet10", "macro_tet4"):
        d[f"{fam}_lap_f64"] = o64.p2tet_laplacian_apply_opt(fam, el, fff, u)
        d[f"{fam}_lap_f32"] = o32.p2tet_laplacian_apply_opt(fam, el, fff, u)
        d[f"{fam}_lap_diag_f64"] = o64.p2tet_laplacian_diag_opt(fam, el, fff, nn)
        d[f"{fam}_le_f64"] = o64.p2tet_linear_elasticity_apply_opt(fam, el, adj, det, 1.3, 0.7, u3)
        d[f"{fam}_le_f32"] = o32.p2tet_linear_elasticity_apply_opt(fam, el, adj, det, 1.3, 0.7, u3)
        d[f"{fam}_le_diag_f64"] = o64.p2tet_linear_elasticity_diag_opt(fam, el, adj, det, 1.3, 0.7, nn)
    np.savez_compressed(os.path.join(OUT, "p2_tet.npz"), **d)
    print("p2_tet.npz", os.path.getsize(os.path.join(OUT, "p2_tet.npz")))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "p2_tet":
        return p2_tet()
    o64, o32 = Oracle("reference", "f64"), Oracle("reference", "f32")
    rng = np.random.default_rng(2026)

    el, pts = mesh.box_tet4(5, 4, 6)
    pts = mesh.warp_points(pts)
    nn = pts.shape[1]
    fff = o64.tet4_fff_fill(el, pts)
    u = rng.uniform(-1, 1, nn)
    np.savez_compressed(os.path.join(OUT, "tet4_laplacian.npz"), elements=el, points=pts, fff=fff, u=u,
                        y_f64=o64.tet4_laplacian_apply_opt(el, fff, u), y_f32=o32.tet4_laplacian_apply_opt(el, fff, u),
                        diag_f64=o64.tet4_laplacian_diag_opt(el, fff, nn))

    el, pts = mesh.box_hex8(5, 4, 4)
    pts = mesh.shear_points(pts)
    nn = pts.shape[1]
    fff = o64.hex8_fff_fill(el, pts)
    adj, det = o64.hex8_adjugate_and_det_fill(el, pts)
    u = rng.uniform(-1, 1, nn)
    u3 = rng.uniform(-1, 1, nn * 3)
    np.savez_compressed(os.path.join(OUT, "hex8.npz"), elements=el, points=pts, fff=fff, adjugate=adj, determinant=det, u=u, u3=u3,
                        mu=1.3, lam=0.7, y_lap_f64=o64.hex8_laplacian_apply_opt(el, fff, u),
                        y_lap_f32=o32.hex8_laplacian_apply_opt(el, fff, u),
                        y_le_f64=o64.hex8_linear_elasticity_apply(el, adj, det, 1.3, 0.7, u3, nn),
                        y_le_f32=o32.hex8_linear_elasticity_apply(el, adj, det, 1.3, 0.7, u3, nn))

    for L in (2, 4, 8):
        m = mesh.sshex8_box(L, 2, 2, 1)
        mp = mesh.shear_points(m["m_points"])
        fff = o64.hex8_fff_fill(m["m_elements"], mp)
        adj, det = o64.hex8_adjugate_and_det_fill(m["m_elements"], mp)
        nn = m["n_nodes"]
        u = rng.uniform(-1, 1, nn)
        d = dict(level=L, elements=m["elements"], m_elements=m["m_elements"], m_points=mp, fff=fff, adjugate=adj,
                 determinant=det, u=u, y_lap_f64=o64.sshex8_laplacian_apply(L, m["elements"], fff, u),
                 y_lap_f32=o32.sshex8_laplacian_apply(L, m["elements"], fff, u), mu=1.3, lam=0.7)
        if L <= 4:
            u3 = rng.uniform(-1, 1, nn * 3)
            d.update(u3=u3, y_le_f64=o64.sshex8_linear_elasticity_apply(L, m["elements"], adj, det, 1.3, 0.7, u3, nn))
        np.savez_compressed(os.path.join(OUT, f"sshex8_L{L}.npz"), **d)
    p2_tet()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

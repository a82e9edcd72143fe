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


def main():
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
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

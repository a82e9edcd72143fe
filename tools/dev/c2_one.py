#!/usr/bin/env python
"""One C2 configuration (hex8 linear elasticity, 128^3), a few applies: the target of ncu captures.
usage: c2_one.py [sfc|gen] [patch|atomic] [E] [f64|f32] [applies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sfem_b200 import lib, mesh  # noqa: E402

order = sys.argv[1] if len(sys.argv) > 1 else "sfc"
scatter = sys.argv[2] if len(sys.argv) > 2 else "patch"
E = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dt = torch.float32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else torch.float64
n = int(sys.argv[5]) if len(sys.argv) > 5 else 3
el, pts = mesh.box_hex8(129, 129, 129)
if order == "sfc":
    el, pts = mesh.sfc_reorder(el, pts)[:2]
nn = pts.shape[1]
adj, det = lib.adjugate_and_det_fill(lib.HEX8, el, pts)
op = lib.linear_elasticity(lib.HEX8, el, adj, det, 1.0, 1.0, nn, dt)
op.set_option("patch_elements", E)
op.set_option("scatter", lib.SCATTER_PATCH if scatter == "patch" else lib.SCATTER_ATOMIC)
x = torch.rand(nn * 3, dtype=dt, device="cuda")
y = torch.zeros_like(x)
for _ in range(n):
    op.apply_overwrite(x, y)
torch.cuda.synchronize()
print("done", order, scatter, E, dt, float(y.norm()))

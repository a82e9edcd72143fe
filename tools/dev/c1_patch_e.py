#!/usr/bin/env python
"""C1 (tet4 Laplacian, 56^3 nodes, 998 250 tets) with 256- and 1024-tet patches (B200FEM_TET4_PATCH_E), L2 flushed between applies."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from sfem_b200 import lib, mesh
from tools.extras_bench import _entry, _time
el, pts = mesh.box_tet4(56, 56, 56)
fff = lib.fff_fill(lib.TET4, el, pts)
nn = pts.shape[1]
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
x = torch.rand(nn, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x); ref = torch.zeros_like(x)
for E in (1024, 256):
    os.environ["B200FEM_TET4_PATCH_E"] = str(E)
    op = lib.laplacian(lib.TET4, el, fff, nn, torch.float64)
    if E == 1024:
        op.set_option("scatter", lib.SCATTER_ATOMIC); op.apply_overwrite(x, ref); op.set_option("scatter", lib.SCATTER_PATCH)
    for fl, name in ((flush, "L2 flushed"), (None, "L2-resident")):
        ms = _time(lambda: op.apply_overwrite(x, y), torch, flush=fl)
        e = _entry(f"C1 tet4 Laplacian f64 56^3 nodes, {E}-tet patches, {name}", nn, op.algorithmic_bytes, ms)
        e["rel_vs_atomic"] = float((y - ref).norm() / ref.norm())
        print(json.dumps(e), flush=True)

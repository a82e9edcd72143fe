#!/usr/bin/env python
"""tet4 Laplacian patch-kernel variants at C4-per-GPU size (260^3 nodes): direct / staged lists / fully pipelined / atomic."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sfem_b200 import lib, mesh  # noqa: E402
from tools.extras_bench import _entry, _time  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 260
el, pts = mesh.box_tet4(n, n, n)
fff = lib.fff_fill(lib.TET4, el, pts)
nn = pts.shape[1]
op = lib.laplacian(lib.TET4, el, fff, nn, torch.float64)
x = torch.rand(nn, dtype=torch.float64, device="cuda")
y = torch.zeros_like(x)
op.set_option("scatter", lib.SCATTER_ATOMIC)
ref = torch.zeros_like(x)
op.apply_overwrite(x, ref)
print(json.dumps(_entry(f"tet4 {n}^3 atomic", nn, op.algorithmic_bytes, _time(lambda: op.apply_overwrite(x, y), torch))), flush=True)
op.set_option("scatter", lib.SCATTER_PATCH)
for v, name in ((2, "direct + staged incidence lists"), (3, "direct + staged lists + element loads before the node gather")):
    op.set_option("patch_pipelined", v)
    op.apply_overwrite(x, y)
    err = float((y - ref).norm() / ref.norm())
    print(json.dumps(_entry(f"tet4 {n}^3 patch {name}", nn, op.algorithmic_bytes, _time(lambda: op.apply_overwrite(x, y), torch), {"rel_vs_atomic": err})), flush=True)

#!/usr/bin/env python
"""C2 (hex8 linear elasticity, 128^3) sweep: scatter strategy x patch size x element order x real type, plus the C5-shard
sshex8 elasticity apply. One JSON line per case (ms, GDOF/s, fraction of the measured HBM peak)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sfem_b200 import lib, mesh  # noqa: E402
from tools.extras_bench import _entry, _time  # noqa: E402

el0, pts0 = mesh.box_hex8(129, 129, 129)
for order in ("generator order", "SFC order"):
    el, pts = (el0, pts0) if order == "generator order" else mesh.sfc_reorder(el0, pts0)[:2]
    nn = pts.shape[1]
    adj, det = lib.adjugate_and_det_fill(lib.HEX8, el, pts)
    for rt, dt in (("f64", torch.float64), ("f32", torch.float32)):
        op = lib.linear_elasticity(lib.HEX8, el, adj, det, 1.0, 1.0, nn, dt)
        x = torch.rand(nn * 3, dtype=dt, device="cuda")
        y = torch.zeros_like(x)
        op.set_option("scatter", lib.SCATTER_ATOMIC)
        ref = torch.zeros_like(x)
        op.apply_overwrite(x, ref)
        ms = _time(lambda: op.apply_overwrite(x, y), torch)
        print(json.dumps(_entry(f"C2 hex8 LE {rt} (atomic, {order})", nn * 3, op.algorithmic_bytes, ms)), flush=True)
        for E in (128, 256, 512):
            op.set_option("patch_elements", E)
            op.set_option("scatter", lib.SCATTER_PATCH)
            op.apply_overwrite(x, y)
            err = float((y - ref).norm() / ref.norm())
            ms = _time(lambda: op.apply_overwrite(x, y), torch)
            print(json.dumps(_entry(f"C2 hex8 LE {rt} (patch E={E}, {order})", nn * 3, op.algorithmic_bytes, ms,
                                    {"rel_vs_atomic": err, "avg_unique_nodes": op.get_option("patch_avg_unique_nodes")})), flush=True)
        del op
from tools.extras_bench import run_extras  # noqa: E402

for e in run_extras(torch, ("c5le",)):
    print(json.dumps(e), flush=True)

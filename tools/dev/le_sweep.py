#!/usr/bin/env python
"""sshex8 linear elasticity at the C5 shard (44^3 macro, L8): tile size sweep (macro elements per CTA)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sfem_b200 import lib, mesh  # noqa: E402
from tools.extras_bench import _entry, _time  # noqa: E402

mm = int(sys.argv[1]) if len(sys.argv) > 1 else 44
m = mesh.sshex8_box(8, mm, mm, mm, with_points=False)
nn = m["n_nodes"]
adj, det = lib.adjugate_and_det_fill(lib.HEX8, m["m_elements"], m["m_points"])
for rt, dt in (("f64", torch.float64), ("f32", torch.float32)):
    op = lib.linear_elasticity(lib.SSHEX8, m["elements"], adj, det, 1.0, 1.0, nn, dt, level=8)
    x = torch.rand(nn * 3, dtype=dt, device="cuda")
    y = torch.zeros_like(x)
    ref = torch.zeros_like(x)
    op.apply_overwrite(x, ref)
    for epb in (6, 3, 2, 1):
        op.set_option("epb", epb)
        op.apply_overwrite(x, y)
        err = float((y - ref).norm() / ref.norm())
        ms = _time(lambda: op.apply_overwrite(x, y), torch, steps=10)
        print(json.dumps(_entry(f"sshex8 LE L8 {rt} {mm}^3 macro, {epb} macro elements per CTA", nn * 3, op.algorithmic_bytes, ms,
                                {"rel_vs_default": err, "fp_ops_TFLOPs": 2.0 * 444 * 512 * m["elements"].shape[1] / ms / 1e9})), flush=True)
    del op, x, y, ref

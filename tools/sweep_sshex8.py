#!/usr/bin/env python
"""Tuning sweep of the sshex8 Laplacian kernel configurations on one GPU (not a bench line)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from sfem_b200 import lib, mesh

ap = argparse.ArgumentParser()
ap.add_argument("--macro", type=int, default=64)
ap.add_argument("--level", type=int, default=8)
ap.add_argument("--cfgs", default="0,1,2,3,4,5")
ap.add_argument("--reals", default="f64,f32")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--check", action="store_true")
ap.add_argument("--tma_cfg", type=int, default=0)
ap.add_argument("--tma", default="0", help="comma list of use_tma values to sweep (0 = plane-streaming, 1 = entity-box TMA kernel)")
ap.add_argument("--ws", default="1", help="comma list of use_ws values (1 = warp-specialised kernel, 0 = plane-streaming kernel)")
ap.add_argument("--ws_cfgs", default="0")
a = ap.parse_args()

m = mesh.sshex8_box(a.level, a.macro, a.macro, a.macro, with_points=False)
fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
nn = m["n_nodes"]
for real in a.reals.split(","):
    dt = torch.float64 if real == "f64" else torch.float32
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, dt, level=a.level)
    x = torch.rand(nn, dtype=dt, device="cuda") - 0.5
    y = torch.zeros_like(x)
    yref = None
    if a.check:
        op.set_option("use_stream", 0)
        yref = torch.zeros_like(x)
        op.apply_overwrite(x, yref)
        op.set_option("use_stream", 1)
    combos = []
    for ws in [int(v) for v in a.ws.split(",")]:
        if ws:
            combos += [(1, 0, c) for c in [int(c) for c in a.ws_cfgs.split(",")]]
        else:
            combos += [(0, t, c) for t in [int(v) for v in a.tma.split(",")] for c in ([int(c) for c in a.cfgs.split(",")] if t == 0 else [0])]
    for ws, tma, cfg in combos:
        op.set_option("use_ws", 2 if ws else 0)
        op.set_option("ws_cfg" if ws else "stream_cfg", cfg)
        op.set_option("use_tma", tma)
        op.set_option("tma_cfg", a.tma_cfg)
        res = {}
        for name, fn in (("overwrite", lambda: op.apply_overwrite(x, y)), ("accumulate", lambda: op.apply(x, y))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            res[name] = ms
        err = None
        if yref is not None:
            op.apply_overwrite(x, y)
            err = float((y - yref).norm() / yref.norm())
        alg = op.algorithmic_bytes
        print(json.dumps({"real": real, "ws": ws, "tma": tma, "cfg": cfg, "ms_overwrite": round(res["overwrite"], 4), "ms_accumulate": round(res["accumulate"], 4),
                          "gdofs_overwrite": round(nn / res["overwrite"] / 1e6, 1), "alg_GBs_overwrite": round(alg / res["overwrite"] / 1e6, 1),
                          "rel_vs_batch_kernel": err}), flush=True)
    del op

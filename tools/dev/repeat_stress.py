#!/usr/bin/env python
"""Repeatability stress: the same kernel on the same input many times; any run that deviates from the first by more than
RED-order rounding is an intermittent error (a race). Small meshes on purpose (the solver tests that flake are small)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from sfem_b200 import lib, mesh

REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dt = torch.float64

def stress(name, fn, ref_scale):
    first = fn().clone()
    worst, bad = 0.0, 0
    for _ in range(REPS):
        out = fn()
        d = float((out - first).abs().max()) / ref_scale
        worst = max(worst, d)
        bad += d > 1e-11
    print(f"{name:44s} reps {REPS} worst rel dev {worst:.2e} deviating runs {bad}", flush=True)

def op_case(name, op, nn, bs=1):
    x = torch.rand(nn * bs, dtype=dt, device="cuda") - 0.5
    y = torch.zeros_like(x)
    def f():
        op.apply_overwrite(x, y)
        return y
    op.apply_overwrite(x, y)
    torch.cuda.synchronize()
    stress(name + " overwrite", f, float(y.abs().max()))
    def g():
        y.zero_()
        op.apply(x, y)
        return y
    stress(name + " accumulate", g, float(y.abs().max()))

for L, dims in ((4, (3, 2, 2)), (8, (3, 2, 2)), (2, (5, 3, 2))):
    m = mesh.sshex8_box(L, *dims, with_points=False)
    fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, m["n_nodes"], dt, level=L)
    op_case(f"sshex8 laplacian L{L}", op, m["n_nodes"])
el, pts = mesh.box_tet4(10, 10, 10)
fff = lib.fff_fill(lib.TET4, el, pts)
op = lib.laplacian(lib.TET4, el, fff, pts.shape[1], dt)
op_case("tet4 laplacian (default scatter)", op, pts.shape[1])
el, pts = mesh.box_hex8(8, 8, 8)
fff = lib.fff_fill(lib.HEX8, el, pts)
op = lib.laplacian(lib.HEX8, el, fff, pts.shape[1], dt)
op_case("hex8 laplacian", op, pts.shape[1])
n = 200003
a = torch.rand(n, dtype=dt, device="cuda"); b = torch.rand(n, dtype=dt, device="cuda")
d0 = lib.blas.dot(a, b); bad = sum(lib.blas.dot(a, b) != d0 for _ in range(REPS)); print("blas dot deviating runs", bad)
n0 = lib.blas.norm2(a); bad = sum(lib.blas.norm2(a) != n0 for _ in range(REPS)); print("blas norm2 deviating runs", bad)

#!/usr/bin/env python
"""CG / multigrid solve times on one GPU through the C++ host layer (sfem_b200/host) -- the solver rows of
BASELINE.json ("CG/MG solve time") at single-GPU size. Not the headline line: bench.py embeds the result under
"solves". Problems follow the reference's drivers:

  cg_poisson   : run_poisson / poisson.exe.cpp pattern on the C3 mesh (sshex8 L8, 64^3 macro, 135 M dofs):
                 u = -1 / +1 on the x-faces, zero rhs, unpreconditioned CG. Unpreconditioned CG needs O(N) iterations
                 at this size, so the timed run is a FIXED number of iterations (max_it), reported per iteration with
                 the HBM roofline of one iteration: B_alg(apply) + 11 vector words per dof.
  mg_elasticity: drivers/mgsolve.cpp, 2 levels: Chebyshev3 (30 sweeps as in notebook/SOLVERS.md:16, eig scale 1.02) on the
                 sshex8 L8 linear-elasticity operator, CG (rtol 1e-8, atol 1e-10) on the derefined hex8 operator,
                 hierarchical restriction / prolongation with zeroed constraints; (-1,0,0) / (1,0,0) on the x-faces.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def _time_gpu(torch, fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def cg_poisson(torch, macro=64, level=8, max_it=200, peak_gbs=6470.5):
    from sfem_b200 import host, lib, mesh

    dt = torch.float64
    m = mesh.sshex8_box(level, macro, macro, macro, with_points=False)
    nn = m["n_nodes"]
    fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, dt, level=level)
    f = host.Function(dt, nn, 1)
    f.add_operator(op)
    f.add_dirichlet(mesh.sshex8_face_nodes(level, m["elements"], m["m_points"], 0, 0.0), 0, -1.0)
    f.add_dirichlet(mesh.sshex8_face_nodes(level, m["elements"], m["m_points"], 0, 1.0), 0, 1.0)
    blas = host.Blas(dt)
    x = torch.zeros(nn, dtype=dt, device="cuda")
    b = torch.zeros(nn, dtype=dt, device="cuda")
    f.apply_constraints(x)
    f.apply_constraints(b)
    cg = host.cg(blas, f, None, rtol=1e-30, atol=1e-30, max_it=max_it)
    host.apply(cg, b, x.clone())  # warm-up (allocates the work vectors)
    rc, ms = _time_gpu(torch, lambda: host.apply(cg, b, x))
    h = host.residual_history(cg)
    it = host.iterations(cg)
    per_it = ms / max(it, 1)
    # per iteration: apply (y = A p) + dot(p, Ap) 2 words + fused x/r/rtr update 6 + p update 3 = 11 words per dof
    bytes_it = op.algorithmic_bytes + nn * 8 * 11
    return {"name": f"CG Poisson, sshex8 L{level} {macro}^3 macro ({nn} dofs), f64, fixed {it} iterations", "dofs": nn,
            "iterations": it, "ms_total": ms, "ms_per_iteration": per_it, "residual_0": float(h[0]), "residual_end": float(h[-1]),
            "alg_GB_s": bytes_it / (per_it * 1e-3) / 1e9, "frac_of_measured_hbm": bytes_it / (per_it * 1e-3) / 1e9 / peak_gbs,
            "gdof_s_per_iteration": nn / (per_it * 1e-3) / 1e9, "vector_kernel_launches": blas.launches()}


def cg_poisson_tet4(torch, n=232, max_it=200, peak_gbs=6470.5):
    """One GPU's share of BASELINE config 4 (~100 M-node tet4 cube over 8 GPUs = 232^3 nodes, 7.4e7 tets per GPU):
    poisson.exe.cpp pattern, u = -1 / +1 on the z-faces, zero rhs, unpreconditioned CG, fixed iteration count."""
    from sfem_b200 import host, lib, mesh

    dt = torch.float64
    t0 = time.time()
    el, pts = mesh.box_tet4(n, n, n)
    nn = pts.shape[1]
    fff = lib.fff_fill(lib.TET4, el, pts)
    op = lib.laplacian(lib.TET4, el, fff, nn, dt)
    f = host.Function(dt, nn, 1)
    f.add_operator(op)
    z = pts[2]
    f.add_dirichlet(np.flatnonzero(z < 1e-6).astype(np.int32), 0, -1.0)
    f.add_dirichlet(np.flatnonzero(z > 1 - 1e-6).astype(np.int32), 0, 1.0)
    setup_s = time.time() - t0
    blas = host.Blas(dt)
    x = torch.zeros(nn, dtype=dt, device="cuda")
    b = torch.zeros(nn, dtype=dt, device="cuda")
    f.apply_constraints(x)
    f.apply_constraints(b)
    cg = host.cg(blas, f, None, rtol=1e-30, atol=1e-30, max_it=max_it)
    host.apply(cg, b, x.clone())
    rc, ms = _time_gpu(torch, lambda: host.apply(cg, b, x))
    h = host.residual_history(cg)
    it = host.iterations(cg)
    per_it = ms / max(it, 1)
    bytes_it = op.algorithmic_bytes + nn * 8 * 11
    return {"name": f"CG Poisson, tet4 {n}^3 nodes ({el.shape[1]} tets; one GPU's share of config 4), f64, fixed {it} iterations",
            "dofs": nn, "iterations": it, "ms_total": ms, "ms_per_iteration": per_it, "residual_0": float(h[0]),
            "residual_end": float(h[-1]), "alg_GB_s": bytes_it / (per_it * 1e-3) / 1e9,
            "frac_of_measured_hbm": bytes_it / (per_it * 1e-3) / 1e9 / peak_gbs, "gdof_s_per_iteration": nn / (per_it * 1e-3) / 1e9,
            "setup_s": setup_s}


def mg_elasticity(torch, macro=32, level=8, sweeps=30, max_cycles=40, tol=1e-9):
    from sfem_b200 import host, lib, mesh

    dt = torch.float64
    t0 = time.time()
    # box edge = macro * level: micro-elements of size 1. The stiffness entries scale with the element size while the
    # reference keeps unscaled identity rows for Dirichlet dofs (Function::apply), so on the unit cube at this
    # resolution the largest eigenvalue IS the identity rows' 1 and the Chebyshev interval misses the operator.
    side = float(macro * level)
    m = mesh.sshex8_box(level, macro, macro, macro, side, side, side, with_points=False)
    nn = m["n_nodes"]
    mel = m["m_elements"]
    n_macro = int(mel.max()) + 1
    adj, det = lib.adjugate_and_det_fill(lib.HEX8, mel, m["m_points"])
    fine_op = lib.linear_elasticity(lib.SSHEX8, m["elements"], adj, det, 1.0, 1.0, nn, dt, level=level)
    coarse_op = lib.linear_elasticity(lib.HEX8, mel, adj, det, 1.0, 1.0, n_macro, dt)
    left = mesh.sshex8_face_nodes(level, m["elements"], m["m_points"], 0, 0.0)
    right = mesh.sshex8_face_nodes(level, m["elements"], m["m_points"], 0, side)
    conds = [(left, 0, -1.0), (left, 1, 0.0), (left, 2, 0.0), (right, 0, 1.0), (right, 1, 0.0), (right, 2, 0.0)]
    f, fc = host.Function(dt, nn, 3), host.Function(dt, n_macro, 3)
    f.add_operator(fine_op)
    fc.add_operator(coarse_op)
    for nodes, comp, val in conds:
        f.add_dirichlet(nodes, comp, val)
        fc.add_dirichlet(nodes[nodes < n_macro], comp, val)
    lat = np.zeros((8, mel.shape[1]), dtype=np.int32)
    for k, (cx, cy, cz) in enumerate(mesh.HEX8_CORNERS):
        lat[mesh.sshex8_lidx(1, cx, cy, cz)] = mel[k]
    tr = lib.Transfer(level, 1, m["elements"], nn, 1, 1, lat, n_macro)
    blas = host.Blas(dt)
    R = host.restriction(dt, tr, 3, fc)
    P = host.prolongation(dt, tr, 3, f)
    setup_s = time.time() - t0
    t0 = time.time()
    # power method from a seeded random vector (Chebyshev3::init_with_random): from the ones vector of mgsolve.cpp:183
    # the interior rows of A 1 vanish and the iteration stops at the eigenvalue 1 of the Dirichlet rows
    guess = torch.from_numpy(np.random.default_rng(5).uniform(-0.5, 0.5, nn * 3)).cuda()
    smoother = host.cheb3(blas, f, eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=sweeps,
                          init=guess)
    del guess
    torch.cuda.synchronize()
    power_s = time.time() - t0
    coarse = host.cg(blas, fc, None, rtol=1e-8, atol=1e-10, max_it=40000)
    mg = host.mg(blas, max_it=max_cycles, atol=tol)
    host.mg_add_level(mg, f, smoother, None, R)
    host.mg_add_level(mg, None, coarse, P, None)
    x = torch.zeros(nn * 3, dtype=dt, device="cuda")
    b = torch.zeros(nn * 3, dtype=dt, device="cuda")
    f.apply_constraints(x)
    f.apply_constraints(b)
    rc, ms = _time_gpu(torch, lambda: host.apply(mg, b, x))
    h = host.residual_history(mg)
    cycles = host.iterations(mg)
    return {"name": f"2-level multigrid (mgsolve), sshex8 L{level} linear elasticity {macro}^3 macro ({nn * 3} dofs), f64, "
                    f"Chebyshev3 x{sweeps} / coarse CG", "dofs": nn * 3, "coarse_dofs": n_macro * 3, "rc": rc, "cycles": cycles,
            "ms_total": ms, "ms_per_cycle": ms / max(cycles, 1), "residual_history": [float(v) for v in h],
            "eig_max": host.cheb3_eig_max(smoother), "power_method_iterations": host.cheb3_power_method_iterations(smoother),
            "power_method_s": power_s, "setup_s": setup_s, "last_coarse_cg_iterations": host.iterations(coarse)}


def run_solves(torch, cg_macro=64, mg_macro=32, peak_gbs=6470.5, tet4_n=232):
    out = []
    cases = [(cg_poisson, dict(macro=cg_macro, peak_gbs=peak_gbs)), (mg_elasticity, dict(macro=mg_macro))]
    if tet4_n:
        cases.append((cg_poisson_tet4, dict(n=tet4_n, peak_gbs=peak_gbs)))
    for fn, kw in cases:
        try:
            out.append(fn(torch, **kw))
        except Exception as ex:  # report, do not hide the headline number
            out.append({"name": fn.__name__, "error": repr(ex)})
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import torch

    cg_macro = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mg_macro = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    tet4_n = int(sys.argv[3]) if len(sys.argv) > 3 else 232
    for r in run_solves(torch, cg_macro, mg_macro, tet4_n=tet4_n):
        print(json.dumps(r), flush=True)

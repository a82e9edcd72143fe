"""The reference DRIVERS' own problems (their real boundary conditions and zero right-hand side, not the seeded-rhs variants of
test_gpu_solvers.py), solved by the reference's solver headers on the CPU (oracle/_ref, live) and by the C++ host layer on the
GPU, side by side:

  * poisson.exe (config 4; drivers/simulations/poisson.exe.cpp:14-41, 122-129): tet4 cube, u = -1 / +1 on the z faces,
    unpreconditioned CG, rtol 1e-6. The answer is the linear field: in exact arithmetic CG terminates as soon as the Krylov
    space spans the (essentially one-dimensional) problem, so WHEN the residual collapses depends on rounding -- the two
    residual histories are compared iteration by iteration up to the collapse, and the counts may differ by a few there.
  * mgsolve (config 5; drivers/mgsolve.cpp:61-131, 170-290, SFEM_USE_CHEB=1): sshex8 L8 linear elasticity, (-1,0,0) / (1,0,0)
    on the x faces, Chebyshev3 x30 + coarse CG; cycle counts must be IDENTICAL and the per-cycle residuals agree.
Both histories go to gpurun_out/driver_problem_histories.json for the record (profiles/)."""
import json
import os

import numpy as np
import pytest

from helpers import backends, problems, solver_cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def both(oracles):
    import torch

    from oracle import ref_solvers

    assert torch.cuda.is_available()
    orc = oracles.get(("reference", "f64"))
    if orc is None or not ref_solvers.available("f64"):
        pytest.skip("needs the reference solver build (oracle/_ref) on this box")
    return orc, backends.RefBackend(orc, ref_solvers.RefSolvers("f64")), backends.GpuBackend("f64")


def _record(name, payload):
    path = os.path.join(ROOT, "gpurun_out", "driver_problem_histories.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[name] = payload
    json.dump(data, open(path, "w"), indent=1)


def test_poisson_exe_problem_cg_histories(both):
    from oracle import ref_solvers
    from sfem_b200 import host

    orc, R, G = both
    p = problems.poisson_tet4(orc, 33)  # the driver's plate problem: unit cube, z faces, zero rhs
    # reference, with its own monitor output captured
    A = R.linear_op(p)
    s = R.cg(A, None, rtol=1e-6, atol=1e-16, max_it=20000)
    R.rs.cg_set_verbose(s, True)
    x, b = solver_cases._bvp_vectors(R, p)
    with ref_solvers.capture_stdout() as cap:
        rc_ref = R.apply(s, b, x)
    h_ref = ref_solvers.parse_cg_monitor(cap.text)
    it_ref = R.iterations(s)
    x_ref = R.to_numpy(x)
    # GPU
    Ag = G.linear_op(p)
    sg = G.cg(Ag, None, rtol=1e-6, atol=1e-16, max_it=20000)
    xg, bg = solver_cases._bvp_vectors(G, p)
    rc_gpu = G.apply(sg, bg, xg)
    h_gpu = host.residual_history(sg)
    it_gpu = G.iterations(sg)
    x_gpu = G.to_numpy(xg)
    ref_abs = np.array([v for _, v in h_ref])
    n = min(len(ref_abs), len(h_gpu))
    rel = np.abs(h_gpu[:n] - ref_abs[:n]) / np.maximum(ref_abs[:n], 1e-300)
    _record("poisson_exe_tet4_33", {"iterations_reference": it_ref, "iterations_gpu": it_gpu, "rc": [rc_ref, rc_gpu],
                                    "history_reference": ref_abs.tolist(), "history_gpu": [float(v) for v in h_gpu],
                                    "max_rel_diff_common_prefix": float(rel.max())})
    print(f"poisson.exe problem: reference {it_ref} iterations, GPU {it_gpu}; max rel diff of the common history {rel.max():.2e}")
    assert rc_ref == 0 and rc_gpu == 0
    # both reach the linear answer
    assert np.abs(x_ref - p.exact).max() < 1e-5 and np.abs(x_gpu - p.exact).max() < 1e-5
    # the histories coincide (printed with 6 digits) while the residual is above the rounding-dominated collapse ...
    smooth = ref_abs[:n] > 1e-3 * ref_abs[0]
    assert rel[smooth].max() < 1e-4, rel[smooth].max()
    # ... and the counts differ by at most the few iterations the collapse takes
    assert abs(it_ref - it_gpu) <= 3, (it_ref, it_gpu)


def test_mgsolve_problem_cycles_identical(both):
    from oracle import ref_solvers
    from sfem_b200 import host, mesh

    orc, R, G = both
    L, mx, my, mz, sweeps = 8, 6, 6, 3, 30

    def make():
        # micro-elements of size 1 (tools/solver_bench.py explains why the unit cube defeats the eigenvalue estimate)
        p = problems.elasticity_sshex8(orc, L, mx, my, mz, scale=float(L))
        return p

    def solve(B, p, verbose=False):
        A, Ac = B.linear_op(p), B.linear_op(p, coarse=True)
        Rr, Pr = B.transfer(p)
        guess = B.from_numpy(np.random.default_rng(5).uniform(-0.5, 0.5, p.n_dofs))
        smoother = B.cheb3(A, eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=sweeps,
                           initial_guess_zero=False, init=guess)
        coarse = B.cg(Ac, None, rtol=1e-8, atol=1e-10, max_it=40000)
        mg = B.mg(max_it=40, atol=1e-9, **({"verbose": True} if verbose else {}))
        B.mg_add_level(mg, A, smoother, None, Rr)
        B.mg_add_level(mg, None, coarse, Pr, None)
        x, b = solver_cases._bvp_vectors(B, p)
        p.ref_restrict_calls = 0
        rc = B.apply(mg, b, x)
        return rc, mg, B.cheb3_eig_max(smoother), B.to_numpy(x)

    p = make()
    with ref_solvers.capture_stdout() as cap:
        rc_ref, mg_ref, eig_ref, x_ref = solve(R, p, verbose=True)
    h_ref = np.array(ref_solvers.parse_mg_monitor(cap.text))
    cycles_ref = p.ref_restrict_calls
    pg = make()
    rc_gpu, mg_gpu, eig_gpu, x_gpu = solve(G, pg)
    h_gpu = host.residual_history(mg_gpu)
    cycles_gpu = G.iterations(mg_gpu)
    _record("mgsolve_elasticity_L8_6x6x3", {"cycles_reference": cycles_ref, "cycles_gpu": cycles_gpu, "eig_max": [eig_ref, eig_gpu],
                                            "history_reference": h_ref.tolist(), "history_gpu": [float(v) for v in h_gpu]})
    print(f"mgsolve problem: reference {cycles_ref} cycles, GPU {cycles_gpu}; eig_max {eig_ref:.10g} / {eig_gpu:.10g}")
    assert rc_ref == 0 and rc_gpu == 0
    assert cycles_ref == cycles_gpu, (cycles_ref, cycles_gpu)
    assert abs(eig_ref - eig_gpu) < 1e-8 * eig_ref
    n = min(len(h_ref), len(h_gpu))
    assert n >= 3 and np.allclose(h_gpu[:n], h_ref[:n], rtol=1e-4)  # printed with 6 digits
    assert np.linalg.norm(x_gpu - x_ref) < 1e-8 * np.linalg.norm(x_ref)

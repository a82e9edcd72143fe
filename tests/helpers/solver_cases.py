"""Solver cases, written once and run on the reference (CPU) and on the B200 host layer (tests/helpers/backends.py).
Each returns {"iterations": int | list, "x": float64 ndarray, ...}. Parameters follow the reference's drivers:
poisson.exe.cpp:20 (rtol 1e-6), frontend/tests/sfem_PoissonTest.cpp:149 (1e-9), drivers/mgsolve.cpp:64-71
(SFEM_TOL 1e-9, SFEM_CHEB_EIG_MAX_SCALE 1.02, SFEM_CHEB_EIG_TOL 1e-5, SFEM_SMOOTHER_SWEEPS 3, coarse CG rtol 1e-8 /
atol 1e-10 / max_it 40000)."""
from __future__ import annotations

import numpy as np


def _bvp_vectors(B, p):
    x = B.zeros(p.n_dofs)
    b = B.from_numpy(p.rhs) if getattr(p, "rhs", None) is not None else B.zeros(p.n_dofs)
    B.apply_constraints(p, x)
    B.apply_constraints(p, b)
    return x, b


def cg(B, p, rtol=1e-6):
    A = B.linear_op(p)
    s = B.cg(A, None, rtol=rtol, atol=1e-16, max_it=10000)
    x, b = _bvp_vectors(B, p)
    rc = B.apply(s, b, x)
    return dict(rc=rc, iterations=B.iterations(s), x=B.to_numpy(x))


def pcg_jacobi(B, p, rtol=1e-9):
    A = B.linear_op(p)
    d = B.diag(p)
    s = B.cg(A, B.jacobi(d, 1.0), rtol=rtol, atol=1e-16, max_it=10000)
    x, b = _bvp_vectors(B, p)
    rc = B.apply(s, b, x)
    return dict(rc=rc, iterations=B.iterations(s), x=B.to_numpy(x))


def stationary_jacobi(B, p, sweeps=25):
    A = B.linear_op(p)
    d = B.diag(p)
    s = B.stationary(A, B.jacobi(d, 0.6), sweeps)
    x, b = _bvp_vectors(B, p)
    rc = B.apply(s, b, x)
    return dict(rc=rc, iterations=B.iterations(s), x=B.to_numpy(x))


def chebyshev(B, p, sweeps=5, random_start=False):
    """power method + one smoothing call on the boundary-value problem. Start vector: ones (mgsolve.cpp:183), or a
    seeded random vector (Chebyshev3::init_with_random) -- from ones the interior rows of A 1 vanish up to rounding,
    so WHEN the dominant eigenvector emerges from the noise depends on the summation order of the kernels."""
    A = B.linear_op(p)
    init = "ones"
    if random_start:
        init = B.from_numpy(np.random.default_rng(5).uniform(-0.5, 0.5, p.n_dofs))
    c = B.cheb3(A, eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=sweeps,
                initial_guess_zero=False, init=init)
    x, b = _bvp_vectors(B, p)
    rc = B.apply(c, b, x)
    return dict(rc=rc, iterations=B.iterations(c), eig_max=B.cheb3_eig_max(c), x=B.to_numpy(x))


def mg_two_level(B, p, sweeps=3, tol=1e-9, max_cycles=40):
    """drivers/mgsolve.cpp: Chebyshev3 smoother on the fine sshex8 operator, CG on the derefined (hex8) operator."""
    A, Ac = B.linear_op(p), B.linear_op(p, coarse=True)
    R, Pr = B.transfer(p)
    smoother = B.cheb3(A, eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=sweeps,
                       initial_guess_zero=False)
    coarse = B.cg(Ac, None, rtol=1e-8, atol=1e-10, max_it=40000)
    mg = B.mg(max_it=max_cycles, atol=tol)
    B.mg_add_level(mg, A, smoother, None, R)
    B.mg_add_level(mg, None, coarse, Pr, None)
    x, b = _bvp_vectors(B, p)
    rc = B.apply(mg, b, x)
    out = dict(rc=rc, eig_max=B.cheb3_eig_max(smoother), x=B.to_numpy(x), last_coarse_iterations=B.iterations(coarse))
    it = B.iterations(mg)  # the reference's Multigrid has no iterations() accessor: -1 there
    # cycles executed = fine-level restrictions performed (sfem_Multigrid.hpp:62-67, 299-303)
    out["iterations"] = it if it >= 0 else p.ref_restrict_calls
    return out


def mg_multilevel(B, p, sweeps=3, tol=1e-9, max_cycles=40):
    """create_ssgmg's N-level hierarchy (levels L, L/2, ..., 1) with mgsolve's Chebyshev3 smoother on every level but
    the coarsest (power method from a seeded random vector) and CG on the coarsest."""
    n = len(p.levels)
    ops = [B.level_op(p, i) for i in range(n)]
    mg = B.mg(max_it=max_cycles, atol=tol)
    smoothers = []
    transfers = [B.level_transfer(p, i) for i in range(n - 1)]
    for i in range(n):
        if i < n - 1:
            guess = B.from_numpy(np.random.default_rng(100 + i).uniform(-0.5, 0.5, p.levels[i]["n_nodes"]))
            sm = B.cheb3(ops[i], eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=sweeps,
                         initial_guess_zero=False, init=guess)
        else:
            sm = B.cg(ops[i], None, rtol=1e-8, atol=1e-10, max_it=40000)
        smoothers.append(sm)
        restriction = transfers[i][0] if i < n - 1 else None
        prolongation = transfers[i - 1][1] if i > 0 else None
        B.mg_add_level(mg, ops[i], sm, prolongation, restriction)
    x, b = _bvp_vectors(B, p)
    p.ref_restrict_calls = 0
    rc = B.apply(mg, b, x)
    out = dict(rc=rc, x=B.to_numpy(x), eig_max=[B.cheb3_eig_max(s) for s in smoothers[:-1]],
               last_coarse_iterations=B.iterations(smoothers[-1]))
    it = B.iterations(mg)
    out["iterations"] = it if it >= 0 else p.ref_restrict_calls
    return out


def pcg_mg(B, p, sweeps=3, rtol=1e-9):
    """SFEM_USE_MG_PRECONDITIONER (drivers/mgsolve.cpp:262-272): CG preconditioned by one multigrid cycle."""
    A, Ac = B.linear_op(p), B.linear_op(p, coarse=True)
    R, Pr = B.transfer(p)
    smoother = B.cheb3(A, eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=sweeps,
                       initial_guess_zero=False)
    coarse = B.cg(Ac, None, rtol=1e-8, atol=1e-10, max_it=40000)
    mg = B.mg(max_it=1, atol=0.0)
    B.mg_add_level(mg, A, smoother, None, R)
    B.mg_add_level(mg, None, coarse, Pr, None)
    ksp = B.cg(A, mg, rtol=rtol, atol=1e-16, max_it=200)
    x, b = _bvp_vectors(B, p)
    rc = B.apply(ksp, b, x)
    return dict(rc=rc, iterations=B.iterations(ksp), x=B.to_numpy(x))


def residual_norm(p, x, apply_fn):
    """|| b - A x || of the constrained operator, evaluated with the oracle on the host."""
    b = np.array(p.rhs, dtype=np.float64) if getattr(p, "rhs", None) is not None else np.zeros(p.n_dofs)
    p.apply_constraints_host(b)
    r = np.zeros(p.n_dofs)
    p.ref_linear_op(apply_fn)(np.asarray(x, dtype=np.float64), r)
    return float(np.linalg.norm(b - r))

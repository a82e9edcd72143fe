"""GPU parity of the solver rows of SURVEY section 8(a) -- ConjugateGradient, PCG + Jacobi, StationaryIteration,
PowerMethod + Chebyshev3, 2-level Multigrid, CG preconditioned by multigrid -- through the C++ host layer
(sfem_b200/host) over the C ABI.

Bars (BASELINE.json north_star): iteration counts IDENTICAL to the reference's own solvers on the same problem;
solutions equal to relative L2 1e-8 in fp64 (both are converged to rtol <= 1e-6, so the difference is the
amplified summation-order rounding of the dot products, not the kernels: those are held to 1e-12 in
test_gpu_operators.py).

Two checks per case: against the committed golden values (tests/golden/solvers.json, from the reference) and, when
oracle/_ref travelled to this box, against a live run of the reference solver. The last test runs the REFERENCE's
solver templates themselves (sfem_cg.hpp, sfem_Multigrid.hpp compiled in oracle/_ref) with a BLAS<T> bound to our
device entry points and our operators: the literal drop-in."""
import json
import os

import numpy as np
import pytest

from helpers import backends, problems, solver_cases

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "solvers.json")))


def _cases(orc):
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_solver_golden", os.path.join(HERE, "golden", "make_solver_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.cases(orc)


@pytest.fixture(scope="module")
def gpu_f64():
    import torch

    assert torch.cuda.is_available()
    return backends.GpuBackend("f64")


@pytest.fixture(scope="module")
def ref_backend(oracles):
    from oracle import ref_solvers

    orc = oracles.get(("reference", "f64"))
    if orc is None or not ref_solvers.available("f64"):
        return None
    return backends.RefBackend(orc, ref_solvers.RefSolvers("f64"))


@pytest.mark.parametrize("name", sorted(GOLD))
def test_solver_matches_reference(oracles, gpu_f64, ref_backend, name):
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    fn, mk, kw = _cases(orc)[name]
    p = mk()
    out = fn(gpu_f64, p, **kw)
    g = GOLD[name]
    assert out["rc"] == g["rc"]
    assert out["iterations"] == g["iterations"], (out["iterations"], g["iterations"])
    if "eig_max" in g:
        assert np.allclose(out["eig_max"], g["eig_max"], rtol=1e-10, atol=0)
    if "last_coarse_iterations" in g:
        assert out["last_coarse_iterations"] == g["last_coarse_iterations"]
    assert abs(np.linalg.norm(out["x"]) - g["x_l2"]) < 1e-8 * g["x_l2"]
    stride = max(1, out["x"].size // 16)
    assert np.allclose(out["x"][::stride][:16], g["x_probe"], rtol=0, atol=1e-8 * g["x_l2"] / np.sqrt(out["x"].size) * 10)
    if ref_backend is not None:
        ref = fn(ref_backend, mk(), **kw)
        assert out["iterations"] == ref["iterations"]
        assert np.linalg.norm(out["x"] - ref["x"]) < 1e-8 * np.linalg.norm(ref["x"])


def test_cg_residual_history_and_known_answer(oracles, gpu_f64):
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    from sfem_b200 import host

    p = problems.poisson_tet4(orc, 10)
    A = gpu_f64.linear_op(p)
    s = gpu_f64.cg(A, None, rtol=1e-10, atol=1e-16, max_it=1000)
    x, b = gpu_f64.zeros(p.n_dofs), gpu_f64.zeros(p.n_dofs)
    gpu_f64.apply_constraints(p, x)
    gpu_f64.apply_constraints(p, b)
    assert gpu_f64.apply(s, b, x) == 0
    h = host.residual_history(s)
    assert len(h) == host.iterations(s) + 2  # initial + one per iteration (the converged one included)
    assert h[-1] / h[0] < 1e-10 <= h[-2] / h[0]
    assert np.abs(gpu_f64.to_numpy(x) - p.exact).max() < 1e-6
    # the true residual agrees with the recurrence
    assert abs(solver_cases.residual_norm(p, gpu_f64.to_numpy(x), p.ref_apply) - h[-1]) < 1e-8 * h[0]


def test_cg_fails_at_max_it_like_the_reference(oracles, gpu_f64):
    """not converged within max_it -> SFEM_FAILURE (algebra/sfem_cg.hpp:328-361)"""
    orc = oracles.get(("reference", "f64")) or oracles[("port", "f64")]
    p = problems.poisson_tet4(orc, 8)
    s = gpu_f64.cg(gpu_f64.linear_op(p), None, rtol=1e-12, atol=1e-30, max_it=3)
    x, b = gpu_f64.zeros(p.n_dofs), gpu_f64.zeros(p.n_dofs)
    gpu_f64.apply_constraints(p, x)
    gpu_f64.apply_constraints(p, b)
    assert gpu_f64.apply(s, b, x) == 1
    assert gpu_f64.iterations(s) == 3
    # zero residual at entry: returns success without iterating (:316-318)
    from sfem_b200 import host, lib
    import torch

    f = host.Function(torch.float64, p.n_nodes, 1)
    f.add_operator(lib.laplacian(lib.TET4, p.elements, p.fff, p.n_nodes, torch.float64))
    s2 = gpu_f64.cg(f, None, rtol=1e-12, atol=1e-30, max_it=3)
    assert gpu_f64.apply(s2, gpu_f64.zeros(p.n_dofs), gpu_f64.zeros(p.n_dofs)) == 0 and gpu_f64.iterations(s2) == 0


def test_cg_f32(oracles):
    """fp32: same recurrences; dot products accumulate in fp64 on the device, in fp32 in the reference's OpenMP_BLAS,
    so iteration counts may differ by rounding: bar = |difference| <= 2 and the fp32 solution bar 1e-5."""
    from oracle import ref_solvers

    orc32 = oracles.get(("reference", "f32"))
    if orc32 is None or not ref_solvers.available("f32"):
        pytest.skip("needs oracle/_ref f32")
    B = backends.GpuBackend("f32")
    R = backends.RefBackend(orc32, ref_solvers.RefSolvers("f32"))
    out = solver_cases.cg(B, problems.poisson_tet4(orc32, 10), rtol=1e-4)
    ref = solver_cases.cg(R, problems.poisson_tet4(orc32, 10), rtol=1e-4)
    assert out["rc"] == 0 and abs(out["iterations"] - ref["iterations"]) <= 2
    assert np.linalg.norm(out["x"] - ref["x"]) < 1e-4 * np.linalg.norm(ref["x"])


def test_reference_solver_templates_drive_our_kernels(oracles):
    """Drop-in: the reference's ConjugateGradient<T>, Chebyshev3<T> and Multigrid<T> (compiled from /root/reference into
    oracle/_ref) run with execution_space = DEVICE, sfem::BLAS<T> bound to b200fem_blas_* and Operator<T>::apply bound
    to our Function (b200fem_op_apply + constraint copy). Iteration counts must equal the CPU reference's."""
    import torch

    from oracle import ref_solvers
    from sfem_b200 import host, lib

    orc = oracles.get(("reference", "f64"))
    if orc is None or not ref_solvers.available("f64"):
        pytest.skip("needs oracle/_ref")
    rs = ref_solvers.RefSolvers("f64")
    rs.set_device_table(lib.raw())
    DEV = ref_solvers.DEVICE
    G = backends.GpuBackend("f64")
    hraw = host.raw()

    def device_op(handle, n):
        return rs.make_op(lambda x, y: hraw.b200host_apply(handle.ptr, x, y), n, n, es=DEV)

    # CG on the tet4 Poisson problem
    p = problems.poisson_tet4(orc, 12, generic=True)
    f = G.function(p)
    cg = rs.cg(device_op(f, p.n_dofs), None, rtol=1e-9, atol=1e-16, max_it=10000, es=DEV)
    x, b = G.zeros(p.n_dofs), G.from_numpy(p.rhs)
    f.apply_constraints(x)
    f.apply_constraints(b)
    assert rs.apply(cg, b.data_ptr(), x.data_ptr()) == 0
    torch.cuda.synchronize()
    assert rs.iterations(cg) == GOLD["cg_tet4_rtol1e-9"]["iterations"]
    assert abs(np.linalg.norm(x.cpu().numpy()) - GOLD["cg_tet4_rtol1e-9"]["x_l2"]) < 1e-8 * GOLD["cg_tet4_rtol1e-9"]["x_l2"]

    # 2-level multigrid with the reference's Multigrid / Chebyshev3 / CG templates on the device
    p = problems.poisson_sshex8(orc, 4, 3, 2, 2, generic=True)
    f, fc = G.function(p), G.function(p, True)
    R, Pr = G.transfer(p)
    nf, nc = p.n_dofs, p.n_coarse_nodes
    A, Ac = device_op(f, nf), device_op(fc, nc)
    Rr = rs.make_op(lambda frm, to: hraw.b200host_apply(R.ptr, frm, to), nc, nf, es=DEV)
    Pp = rs.make_op(lambda frm, to: hraw.b200host_apply(Pr.ptr, frm, to), nf, nc, es=DEV)
    cheb = rs.cheb3(A, eig_tol=1e-5, eig_max_it=1000, scale_eig_max=1.02, scale_eig_min=0.06, max_it=3, es=DEV)
    assert abs(rs.cheb3_eig_max(cheb) - GOLD["mg2_poisson_L4"]["eig_max"]) < 1e-10
    coarse = rs.cg(Ac, None, rtol=1e-8, atol=1e-10, max_it=40000, es=DEV)
    mg = rs.mg(max_it=40, atol=1e-9, es=DEV)
    rs.mg_add_level(mg, A, cheb, None, Rr)
    rs.mg_add_level(mg, None, coarse, Pp, None)
    x, b = G.zeros(nf), G.from_numpy(p.rhs)
    f.apply_constraints(x)
    f.apply_constraints(b)
    assert rs.apply(mg, b.data_ptr(), x.data_ptr()) == 0
    torch.cuda.synchronize()
    assert abs(np.linalg.norm(x.cpu().numpy()) - GOLD["mg2_poisson_L4"]["x_l2"]) < 1e-8 * GOLD["mg2_poisson_L4"]["x_l2"]

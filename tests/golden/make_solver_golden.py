"""Generates tests/golden/solvers.json: iteration counts, eigenvalue estimates and solution checksums of the
REFERENCE's own solvers (oracle/_ref/libsfem_ref_solvers_f64.so = algebra/sfem_cg.hpp, sfem_Chebyshev3.hpp,
sfem_Multigrid.hpp ... compiled from /root/reference) on the problems of tests/helpers/problems.py.
Run where /root/reference exists:  python tests/golden/make_solver_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from helpers import backends, problems, solver_cases  # noqa: E402
from oracle.ref import Oracle  # noqa: E402
from oracle.ref_solvers import RefSolvers  # noqa: E402


def cases(orc):
    return {
        "cg_tet4_rtol1e-6": (solver_cases.cg, lambda: problems.poisson_tet4(orc, 12, generic=True), dict(rtol=1e-6)),
        "cg_tet4_rtol1e-9": (solver_cases.cg, lambda: problems.poisson_tet4(orc, 12, generic=True), dict(rtol=1e-9)),
        "pcg_jacobi_tet4": (solver_cases.pcg_jacobi, lambda: problems.poisson_tet4(orc, 12, generic=True), {}),
        "stationary_jacobi_tet4": (solver_cases.stationary_jacobi, lambda: problems.poisson_tet4(orc, 8, generic=True), {}),
        "cg_sshex8_L4": (solver_cases.cg, lambda: problems.poisson_sshex8(orc, 4, 3, 2, 2, generic=True), dict(rtol=1e-9)),
        "cheb_sshex8_L4": (solver_cases.chebyshev, lambda: problems.poisson_sshex8(orc, 4, 3, 2, 2, generic=True), {}),
        "cheb_elasticity_L2": (solver_cases.chebyshev, lambda: problems.elasticity_sshex8(orc, 2, 3, 2, 2, generic=True),
                               dict(random_start=True)),
        "mg2_poisson_L4": (solver_cases.mg_two_level, lambda: problems.poisson_sshex8(orc, 4, 3, 2, 2, generic=True), {}),
        "mg2_poisson_L8": (solver_cases.mg_two_level, lambda: problems.poisson_sshex8(orc, 8, 3, 2, 2, generic=True), dict(sweeps=10)),
        "mg2_elasticity_L4": (solver_cases.mg_two_level, lambda: problems.elasticity_sshex8(orc, 4, 3, 2, 2, generic=True), dict(sweeps=8, max_cycles=60)),
        "mg4_poisson_L8": (solver_cases.mg_multilevel, lambda: problems.poisson_sshex8_multilevel(orc, 8, 2, 2, 2), dict(sweeps=3)),
        "mg3_poisson_L4": (solver_cases.mg_multilevel, lambda: problems.poisson_sshex8_multilevel(orc, 4, 3, 2, 2), dict(sweeps=3)),
        "pcg_mg_poisson_L4": (solver_cases.pcg_mg, lambda: problems.poisson_sshex8(orc, 4, 3, 2, 2, generic=True), {}),
    }


def summarise(out):
    x = out.pop("x")
    out["x_l2"] = float(np.linalg.norm(x))
    out["x_sum"] = float(x.sum())
    out["x_probe"] = [float(v) for v in x[:: max(1, x.size // 16)][:16]]
    return out


if __name__ == "__main__":
    orc, rs = Oracle("reference", "f64"), RefSolvers("f64")
    B = backends.RefBackend(orc, rs)
    gold = {}
    for name, (fn, mk, kw) in cases(orc).items():
        gold[name] = summarise(fn(B, mk(), **kw))
        print(name, {k: v for k, v in gold[name].items() if k != "x_probe"}, flush=True)
    with open(os.path.join(HERE, "solvers.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)

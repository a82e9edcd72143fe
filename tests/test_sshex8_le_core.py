"""The per-micro-element arithmetic of the sshex8 linear-elasticity CUDA kernel
(sfem_b200/csrc/hex8_le_core.cuh: gradient moments in the Haar basis instead of the reference's 24x24
element-matrix product) run sequentially on the host -- the same header, compiled with g++ -- against
the oracle's affine_sshex8_linear_elasticity_apply_macro_adjugate. Bars: 1e-12 rel. L2 (fp64), 1e-5 (fp32)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def core_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("lecore") / "libsshex8_le_core_host.so"
    src = os.path.join(ROOT, "tests", "helpers", "sshex8_le_core_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "sfem_b200", "csrc"),
                    "-o", str(out), src], check=True)
    return C.CDLL(str(out))


def _setup(orc, level):
    from sfem_b200 import mesh

    m = mesh.sshex8_box(level, 2, 2, 3)
    mpts = mesh.shear_points(m["m_points"])
    adj, det = orc.hex8_adjugate_and_det_fill(m["m_elements"], mpts)
    return m, mpts, adj, det


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 3, 4, 8])
def test_le_core_matches_oracle(core_lib, oracles, real, level):
    from oracle.ref import _soa_ptrs

    orc = oracles.get(("reference", real)) or oracles[("port", real)]
    m, _, adj, det = _setup(orc, level)
    nn = m["n_nodes"]
    mu, lam = 1.3, 0.7
    u = np.random.default_rng(level).uniform(-1, 1, nn * 3).astype(orc.rt)
    ref = orc.sshex8_linear_elasticity_apply(level, m["elements"], adj, det, mu, lam, u, nn)
    out = np.zeros(nn * 3, dtype=orc.rt)
    el = np.ascontiguousarray(m["elements"])
    fn = getattr(core_lib, "sshex8_le_core_apply_" + real)
    fn.restype = C.c_int
    rc = fn(C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), adj.ctypes.data_as(C.POINTER(C.c_float)),
            det.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(mu), C.c_double(lam), u.ctypes.data_as(C.c_void_p),
            out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    err = np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref)
    assert err < (1e-12 if real == "f64" else 1e-5), err


@pytest.mark.parametrize("level", [2, 4])
def test_le_core_block_diag_matches_oracle(core_lib, oracles, level):
    from oracle.ref import _soa_ptrs

    orc = oracles.get(("reference", "f64"))
    if orc is None:
        pytest.skip("block diagonal needs the reference build (oracle/_ref)")
    m, mpts, adj, det = _setup(orc, level)
    nn = m["n_nodes"]
    mu, lam = 1.3, 0.7
    pts = np.zeros((3, nn), dtype=np.float32)
    # the reference reads the 8 macro corners from `points` indexed by the sshex8 node ids
    from sfem_b200 import mesh

    corners = [mesh.sshex8_lidx(level, cx * level, cy * level, cz * level) for cx, cy, cz in mesh.HEX8_CORNERS]
    for k, c in enumerate(corners):
        pts[:, m["elements"][c]] = mpts[:, m["m_elements"][k]]
    ref6 = orc.sshex8_linear_elasticity_block_diag_sym(level, m["elements"], pts, mu, lam, nn)
    refd = orc.sshex8_linear_elasticity_diag(level, m["elements"], pts, mu, lam, nn)
    out6 = np.zeros(nn * 6)
    el = np.ascontiguousarray(m["elements"])
    fn = core_lib.sshex8_le_core_block_diag_f64
    fn.restype = C.c_int
    fn(C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), adj.ctypes.data_as(C.POINTER(C.c_float)),
       det.ctypes.data_as(C.POINTER(C.c_float)), C.c_double(mu), C.c_double(lam), out6.ctypes.data_as(C.c_void_p))
    # geometry is stored in fp32 on our side and recomputed from fp32 points in fp64 by the reference
    assert np.linalg.norm(out6 - ref6) / np.linalg.norm(ref6) < 1e-6
    d = out6.reshape(nn, 6)[:, [0, 3, 5]].reshape(-1)
    assert np.linalg.norm(d - refd) / np.linalg.norm(refd) < 1e-6

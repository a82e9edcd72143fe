"""The per-thread arithmetic of the sshex8 Laplacian CUDA kernel (sfem_b200/csrc/sshex8_laplacian_core.cuh)
run sequentially on the host -- the same header, compiled with g++ -- against the oracle. This pins the
formulation (one coefficient set per macro element instead of the reference's 27 boundary-class stencils)
without a GPU. Bars: 1e-12 relative L2 in fp64, 1e-5 in fp32."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def core_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("core") / "libsshex8_core_host.so"
    src = os.path.join(ROOT, "tests", "helpers", "sshex8_core_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "sfem_b200", "csrc"),
                    "-o", str(out), src], check=True)
    return C.CDLL(str(out))


@pytest.mark.parametrize("real", ["f64", "f32"])
@pytest.mark.parametrize("level", [2, 3, 4, 8])
def test_core_matches_oracle(core_lib, oracles, real, level):
    from oracle.ref import _soa_ptrs
    from sfem_b200 import mesh

    orc = oracles.get(("reference", real)) or oracles[("port", real)]
    m = mesh.sshex8_box(level, 3, 2, 2)
    fff = orc.hex8_fff_fill(m["m_elements"], mesh.shear_points(m["m_points"]))
    nn = m["n_nodes"]
    u = np.random.default_rng(level).uniform(-1, 1, nn).astype(orc.rt)
    ref = orc.sshex8_laplacian_apply(level, m["elements"], fff, u)
    out = np.zeros(nn, dtype=orc.rt)
    el = np.ascontiguousarray(m["elements"])
    fn = getattr(core_lib, "sshex8_core_apply_" + real)
    fn.restype = C.c_int
    rc = fn(C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), fff.ctypes.data_as(C.POINTER(C.c_float)),
            u.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    err = np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref)
    assert err < (1e-12 if real == "f64" else 1e-5), err
    # constants are in the kernel
    out[:] = 0
    ones = np.ones(nn, dtype=orc.rt)
    fn(C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), fff.ctypes.data_as(C.POINTER(C.c_float)),
       ones.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.abs(out).max() < (1e-12 if real == "f64" else 1e-4)

"""ctypes access to the oracle libraries. TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this module; nothing under `sfem_b200/` does.

Two interchangeable back-ends expose the same extern "C" entry points (the reference's own
operator-level C functions, e.g. `tet4_laplacian_apply_opt`,
operators/tet4/tet4_laplacian.cpp:237 of the reference):

* kind="reference": `oracle/_ref/libsfem_ref_{f64,f32}[_fast].so` -- the reference's own sources
  compiled out of tree by `oracle/Makefile` (target `ref`);
* kind="port": `oracle/liboracle_port_{f64,f32}.so` -- our plain-C restatement
  (`oracle/oracle_port.c`), pinned against the reference build by `tests/test_oracle.py` and
  against the committed fixtures in `tests/golden/`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

c_idx_pp = C.POINTER(C.POINTER(C.c_int32))
c_f32_pp = C.POINTER(C.POINTER(C.c_float))


def _soa_ptrs(arr: np.ndarray, ctype):
    """pointer-to-pointers view of a C-contiguous [k][n] array (keeps `arr` alive via closure)."""
    assert arr.flags["C_CONTIGUOUS"]
    k = arr.shape[0]
    ptrs = (C.POINTER(ctype) * k)()
    for i in range(k):
        ptrs[i] = arr[i].ctypes.data_as(C.POINTER(ctype))
    return ptrs


def lib_path(kind: str, real: str, fast=False) -> str:
    """fast: False = the -O2 parity build, True = the reference's Release flags with -mavx2 -mfma, "avx512" = the same
    with -march=x86-64-v4 (the reference builds -march=native; the GPU box's ISA is not known when this is compiled)."""
    if kind == "reference":
        return os.path.join(HERE, "_ref", f"libsfem_ref_{real}{('_fast_avx512' if fast == 'avx512' else '_fast') if fast else ''}.so")
    if kind == "port":
        return os.path.join(HERE, f"liboracle_port_{real}.so")
    raise ValueError(kind)


def available(kind: str, real: str = "f64", fast=False) -> bool:
    return os.path.exists(lib_path(kind, real, fast))


class Oracle:
    """Thin numpy front for one oracle library (one real type)."""

    def __init__(self, kind: str = "reference", real: str = "f64", fast=False):
        self.kind, self.real = kind, real
        self.rt = np.float64 if real == "f64" else np.float32
        self.crt = C.c_double if real == "f64" else C.c_float
        path = lib_path(kind, real, fast)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle {'ref' if kind == 'reference' else 'port'}`")
        self.lib = C.CDLL(path)

    # ---- helpers
    def _r(self, a):
        a = np.ascontiguousarray(a, dtype=self.rt)
        return a, a.ctypes.data_as(C.POINTER(self.crt))

    @staticmethod
    def _f32(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        return a, a.ctypes.data_as(C.POINTER(C.c_float))

    def _symbol(self, name):
        """C symbol, or the Itanium-mangled C++ one (a few reference headers carry no extern "C": tet4_linear_elasticity.hpp)"""
        try:
            return getattr(self.lib, name)
        except AttributeError:
            pass
        if not hasattr(self, "_mangled"):
            import subprocess

            out = subprocess.run(["nm", "-D", "--defined-only", self.lib._name], capture_output=True, text=True).stdout
            self._mangled = [ln.split()[-1] for ln in out.splitlines() if ln.split()[-1].startswith("_Z")]
        prefix = f"_Z{len(name)}{name}"
        hits = [m for m in self._mangled if m.startswith(prefix) and not m[len(prefix)].isalnum() or m.startswith(prefix + "l")]
        hits = [m for m in self._mangled if m.startswith(prefix) and (len(m) == len(prefix) or not m[len(prefix)].islower() or m[len(prefix)] in "lv")]
        if len(hits) != 1:
            raise AttributeError(f"{name}: {len(hits)} mangled candidates in {self.lib._name}")
        return getattr(self.lib, hits[0])

    def _call(self, name, *args):
        fn = self._symbol(name)
        fn.restype = C.c_int
        rc = fn(*args)
        if rc != 0:
            raise RuntimeError(f"{name} returned {rc}")

    # ---- geometry
    def tet4_fff_fill(self, elements, points):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        ne = el.shape[1]
        fff = np.zeros(ne * 6, dtype=np.float32)
        self._call("tet4_fff_fill", C.c_ssize_t(ne), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float),
                   fff.ctypes.data_as(C.POINTER(C.c_float)))
        return fff

    def hex8_fff_fill(self, elements, points):
        """FFF at the element centre from the 8 corners (works for the macro corners of sshex8)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        ne = el.shape[1]
        fff = np.zeros(ne * 6, dtype=np.float32)
        self._call("hex8_fff_fill", C.c_ssize_t(ne), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float),
                   fff.ctypes.data_as(C.POINTER(C.c_float)))
        return fff

    def hex8_adjugate_and_det_fill(self, elements, points):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        ne = el.shape[1]
        adj = np.zeros(ne * 9, dtype=np.float32)
        det = np.zeros(ne, dtype=np.float32)
        self._call("hex8_adjugate_and_det_fill", C.c_ssize_t(ne), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float),
                   adj.ctypes.data_as(C.POINTER(C.c_float)), det.ctypes.data_as(C.POINTER(C.c_float)))
        return adj, det

    # ---- Laplacian
    def tet4_laplacian_apply_opt(self, elements, fff, u, out=None):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        fff, pf = self._f32(fff)
        u, pu = self._r(u)
        out = np.zeros_like(u) if out is None else np.ascontiguousarray(out, dtype=self.rt)
        self._call("tet4_laplacian_apply_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf, pu,
                   out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def tet4_laplacian_diag_opt(self, elements, fff, nnodes):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        fff, pf = self._f32(fff)
        out = np.zeros(nnodes, dtype=self.rt)
        self._call("tet4_laplacian_diag_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf,
                   out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def hex8_laplacian_apply_opt(self, elements, fff, u, out=None):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        fff, pf = self._f32(fff)
        u, pu = self._r(u)
        out = np.zeros_like(u) if out is None else np.ascontiguousarray(out, dtype=self.rt)
        self._call("hex8_laplacian_apply_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf, pu,
                   out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def sshex8_laplacian_apply(self, level, elements, fff, u, out=None):
        """affine_sshex8_laplacian_stencil_apply_fff (operators/sshex8/sshex8_laplacian.cpp:698)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        fff, pf = self._f32(fff)
        u, pu = self._r(u)
        out = np.zeros_like(u) if out is None else np.ascontiguousarray(out, dtype=self.rt)
        self._call("affine_sshex8_laplacian_stencil_apply_fff", C.c_int(level), C.c_ssize_t(el.shape[1]),
                   _soa_ptrs(el, C.c_int32), pf, pu, out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    # ---- linear elasticity (u/out AoS [nn*3])
    def hex8_linear_elasticity_apply(self, elements, adjugate_aos, det, mu, lam, u, nnodes, out=None):
        """affine_hex8_linear_elasticity_apply_adjugate_soa (operators/hex8/hex8_linear_elasticity.cpp:625);
        the Op hands it SoA adjugate streams (frontend/ops/sfem_LinearElasticity.cpp:275-277)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        ne = el.shape[1]
        adj_soa = np.ascontiguousarray(np.asarray(adjugate_aos, dtype=np.float32).reshape(ne, 9).T)
        det, pd = self._f32(det)
        u, _ = self._r(u)
        out = np.zeros_like(u) if out is None else np.ascontiguousarray(out, dtype=self.rt)
        sz = u.itemsize
        up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("affine_hex8_linear_elasticity_apply_adjugate_soa", C.c_ssize_t(ne), C.c_ssize_t(nnodes),
                   _soa_ptrs(el, C.c_int32), _soa_ptrs(adj_soa, C.c_float), pd, self.crt(mu), self.crt(lam),
                   C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    def sshex8_linear_elasticity_apply(self, level, elements, adjugate_aos, det, mu, lam, u, nnodes, out=None):
        """affine_sshex8_linear_elasticity_apply_macro_adjugate
        (operators/sshex8/sshex8_linear_elasticity.cpp:275)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        ne = el.shape[1]
        adj, pa = self._f32(adjugate_aos)
        det, pd = self._f32(det)
        u, _ = self._r(u)
        out = np.zeros_like(u) if out is None else np.ascontiguousarray(out, dtype=self.rt)
        sz = u.itemsize
        up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("affine_sshex8_linear_elasticity_apply_macro_adjugate", C.c_int(level), C.c_ssize_t(ne),
                   C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32), None, pa, pd, self.crt(mu), self.crt(lam),
                   C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    # ---- non-default paths that take the points (operators/tet4/tet4_laplacian.cpp:70, hex8_linear_elasticity.cpp:86,
    #      sshex8_laplacian.cpp:39 / :184) and the stored-element-matrix operators
    def tet4_laplacian_apply_points(self, elements, points, u, nnodes):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        u, pu = self._r(u)
        out = np.zeros_like(u)
        self._call("tet4_laplacian_apply", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), pu, out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def hex8_vector_laplacian_apply(self, elements, fff, vector_size, u, nnodes):
        """affine_hex8_vector_laplacian_apply_fff (operators/hex8/hex8_vector_laplacian.cpp:78-138); u / out AoS [nn * vs]."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        f, pf = self._f32(fff)
        u, _ = self._r(u)
        out = np.zeros_like(u)
        sz = u.itemsize
        pp = C.POINTER(self.crt) * vector_size
        up = pp(*[C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(vector_size)])
        op = pp(*[C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(vector_size)])
        self._call("affine_hex8_vector_laplacian_apply_fff", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf,
                   C.c_int(vector_size), C.c_ssize_t(vector_size), up, op)
        return out

    def sshex8_vector_laplacian_apply(self, level, elements, fff, vector_size, u):
        """affine_sshex8_vector_laplacian_apply_fff (operators/sshex8/sshex8_vector_laplacian.cpp:16-97); u / out AoS [nn * vs]."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        f, pf = self._f32(fff)
        u, _ = self._r(u)
        out = np.zeros_like(u)
        sz = u.itemsize
        pp = C.POINTER(self.crt) * vector_size
        up = pp(*[C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(vector_size)])
        op = pp(*[C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(vector_size)])
        self._call("affine_sshex8_vector_laplacian_apply_fff", C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf,
                   C.c_int(vector_size), C.c_ssize_t(vector_size), up, op)
        return out

    def sshex8_lumped_mass(self, level, elements, points, nnodes):
        """affine_sshex8_mass_lumped (operators/sshex8/sshex8_mass.cpp:11-115; C++ linkage): points read at the macro corners"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes, dtype=self.rt)
        self._call("affine_sshex8_mass_lumped", C.c_int(level), C.c_ssize_t(el.shape[1]), C.c_ssize_t(0), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    # ---- mass (reference build only: hex8_mass.cpp / tet4_mass.cpp compiled where they lie)
    def hex8_apply_mass(self, elements, points, u, nnodes):
        """hex8_apply_mass (operators/hex8/hex8_mass.cpp:8-60): out += M u, 27-point Gauss, isoparametric."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        u, pu = self._r(u)
        out = np.zeros_like(u)
        self._call("hex8_apply_mass", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), C.c_ssize_t(1), pu, C.c_ssize_t(1), out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def hex8_lumped_mass(self, elements, points, nnodes):
        """hex8_assemble_lumped_mass (operators/hex8/hex8_mass.cpp:62-103)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes, dtype=self.rt)
        self._call("hex8_assemble_lumped_mass", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), C.c_ssize_t(1), out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def tet4_lumped_mass(self, elements, points, nnodes):
        """tet4_assemble_lumped_mass (operators/tet4/tet4_mass.cpp:231-282; C++ linkage, void)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes, dtype=self.rt)
        fn = self._symbol("tet4_assemble_lumped_mass")
        fn.restype = None
        fn(C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float),
           out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def tet10_lumped_mass(self, elements, points, nnodes):
        """tet10_assemble_lumped_mass (operators/tet10/tet10_mass.cpp:528-583; C++ linkage, void, prints its timing)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes, dtype=self.rt)
        fn = self._symbol("tet10_assemble_lumped_mass")
        fn.restype = None
        fn(C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float),
           out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def hex8_linear_elasticity_apply_isoparametric(self, elements, points, mu, lam, u, nnodes):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        u, _ = self._r(u)
        out = np.zeros_like(u)
        sz = u.itemsize
        up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("hex8_linear_elasticity_apply", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3),
                   op[0], op[1], op[2])
        return out

    def sshex8_laplacian_apply_points(self, level, elements, points, u, affine=False):
        """sshex8_laplacian_apply (non-affine, q6) or affine_sshex8_laplacian_apply (element-matrix formulation);
        `points` is SoA fp32 [3][nnodes] by sshex8 node id (only the macro corners are read)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        u, pu = self._r(u)
        out = np.zeros_like(u)
        self._call("affine_sshex8_laplacian_apply" if affine else "sshex8_laplacian_apply", C.c_int(level), C.c_ssize_t(el.shape[1]),
                   _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), pu, out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def sshex8_laplacian_element_matrix(self, level, macro_elements, macro_points):
        """[ne, 64] (sshex8_laplacian_element_matrix, sshex8_laplacian.cpp:595); scalar_t = real_t in this build"""
        el = np.ascontiguousarray(macro_elements, dtype=np.int32)
        pts = np.ascontiguousarray(macro_points, dtype=np.float32)
        out = np.zeros((el.shape[1], 64), dtype=self.rt)
        self._call("sshex8_laplacian_element_matrix", C.c_int(level), C.c_ssize_t(el.shape[1]), C.c_ssize_t(pts.shape[1]),
                   _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def sshex8_linear_elasticity_element_matrix(self, level, macro_elements, macro_points, mu, lam):
        el = np.ascontiguousarray(macro_elements, dtype=np.int32)
        pts = np.ascontiguousarray(macro_points, dtype=np.float32)
        out = np.zeros((el.shape[1], 576), dtype=self.rt)
        self._call("sshex8_linear_elasticity_element_matrix", C.c_int(level), C.c_ssize_t(el.shape[1]), C.c_ssize_t(pts.shape[1]),
                   _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam),
                   out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def sshex8_em_apply(self, level, elements, element_matrix, u, block_size=1):
        """sshex8_stencil_element_matrix_apply / _apply3 (operators/stencil/sshex8_stencil_element_matrix_apply.cpp)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        m = np.ascontiguousarray(element_matrix, dtype=self.rt)
        u, pu = self._r(u)
        out = np.zeros_like(u)
        if block_size == 1:
            self._call("sshex8_stencil_element_matrix_apply", C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32),
                       m.ctypes.data_as(C.POINTER(self.crt)), pu, out.ctypes.data_as(C.POINTER(self.crt)))
            return out
        sz = u.itemsize
        up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("sshex8_stencil_element_matrix_apply3", C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32),
                   m.ctypes.data_as(C.POINTER(self.crt)), C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    def tet4_linear_elasticity_apply_points(self, elements, points, mu, lam, u, nnodes):
        """tet4_linear_elasticity_apply (operators/tet4/tet4_linear_elasticity.cpp:86-160): adjugate from the points in scalar_t"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        u, _ = self._r(u)
        out = np.zeros_like(u)
        sz = u.itemsize
        up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("tet4_linear_elasticity_apply", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3),
                   op[0], op[1], op[2])
        return out

    def tet4_linear_elasticity_diag_points(self, elements, points, mu, lam, nnodes):
        """tet4_linear_elasticity_diag (:162-230)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 3, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("tet4_linear_elasticity_diag", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    # ---- 10-node tetrahedra (reference build only): P2 (tet10) and P1-iso-P2 (macro_tet4), geometry of the 4 corners
    def p2tet_laplacian_apply_opt(self, family, elements, fff, u):
        """tet10_laplacian_apply_opt (operators/tet10/tet10_laplacian.cpp:232-266; C++ linkage) /
        macro_tet4_laplacian_apply_opt (operators/macro_tet4/macro_tet4_laplacian.cpp:401-440)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        fff, pf = self._f32(fff)
        u, pu = self._r(u)
        out = np.zeros_like(u)
        self._call(f"{family}_laplacian_apply_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf, pu,
                   out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def p2tet_laplacian_diag_opt(self, family, elements, fff, nnodes):
        """tet10_laplacian_diag_opt (tet10_laplacian.cpp:268-302) / macro_tet4_laplacian_diag_opt (macro_tet4_laplacian.cpp:442-490)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        fff, pf = self._f32(fff)
        out = np.zeros(nnodes, dtype=self.rt)
        self._call(f"{family}_laplacian_diag_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pf,
                   out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def p2tet_linear_elasticity_apply_opt(self, family, elements, adjugate_aos, det, mu, lam, u):
        """tet10_linear_elasticity_apply_opt (operators/tet10/tet10_linear_elasticity.cpp:209-283) /
        macro_tet4_linear_elasticity_apply_opt (operators/macro_tet4/macro_tet4_linear_elasticity.cpp:206-277)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        adj, pa = self._f32(adjugate_aos)
        det, pd = self._f32(det)
        u, _ = self._r(u)
        out = np.zeros_like(u)
        sz = u.itemsize
        up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call(f"{family}_linear_elasticity_apply_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pa, pd,
                   self.crt(mu), self.crt(lam), C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    def p2tet_laplacian_points(self, family, elements, points, u, nnodes, diag=False):
        """tet10_laplacian_apply / _diag (operators/tet10/tet10_laplacian.cpp:65-230), macro_tet4_laplacian_apply / _diag
        (operators/macro_tet4/macro_tet4_laplacian.cpp:127-260): FFF from the points in scalar_t"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes, dtype=self.rt)
        po = out.ctypes.data_as(C.POINTER(self.crt))
        if diag:
            self._call(f"{family}_laplacian_diag", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                       _soa_ptrs(pts, C.c_float), po)
        else:
            u, pu = self._r(u)
            self._call(f"{family}_laplacian_apply", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                       _soa_ptrs(pts, C.c_float), pu, po)
        return out

    def p2tet_linear_elasticity_points(self, family, elements, points, mu, lam, u, nnodes, diag=False):
        """tet10_linear_elasticity_apply / _diag (operators/tet10/tet10_linear_elasticity.cpp:21-207), macro_tet4_linear_elasticity_apply
        / _diag (operators/macro_tet4/macro_tet4_linear_elasticity.cpp:137-204, 330-395)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 3, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        head = (C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam))
        if diag:
            self._call(f"{family}_linear_elasticity_diag", *head, C.c_ssize_t(3), op[0], op[1], op[2])
        else:
            u, _ = self._r(u)
            up = [C.cast(u.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
            self._call(f"{family}_linear_elasticity_apply", *head, C.c_ssize_t(3), up[0], up[1], up[2], C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    def p2tet_linear_elasticity_diag_opt(self, family, elements, adjugate_aos, det, mu, lam, nnodes):
        """tet10_linear_elasticity_diag_opt (tet10_linear_elasticity.cpp:285-355) / macro_tet4_linear_elasticity_diag_opt
        (macro_tet4_linear_elasticity.cpp:397-437)"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        adj, pa = self._f32(adjugate_aos)
        det, pd = self._f32(det)
        out = np.zeros(nnodes * 3, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call(f"{family}_linear_elasticity_diag_opt", C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32), pa, pd,
                   self.crt(mu), self.crt(lam), C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    def tet4_linear_elasticity_block_diag_sym(self, elements, points, mu, lam, nnodes):
        """tet4_linear_elasticity_block_diag_sym (operators/tet4/tet4_linear_elasticity.cpp:360-440; C++ linkage): [nn * 6] AoS,
        entries (0,0) (0,1) (0,2) (1,1) (1,2) (2,2) of the 3 x 3 node block"""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 6, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(6)]
        self._call("tet4_linear_elasticity_block_diag_sym", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(6), *op)
        return out

    def packed_laplacian_apply(self, element_type, pm, fff, u, two_pass=False):
        """packed_laplacian_apply_u16 of oracle_port.c (port build only): the pack loop of PackedLaplacianApply::apply / ::apply_two_pass
        (frontend/ops/sfem_PackedLaplacian.cpp:111-262) restated -- that file needs smesh and cannot be compiled into oracle/_ref.
        `pm` carries the PackedMesh arrays (sfem_b200.lib.PackedMesh); element_type is the smesh::ElemType value (4, 8, 10, 100)."""
        el = np.ascontiguousarray(pm.elements, dtype=np.uint16)
        f, pf = self._f32(fff)
        u, pu = self._r(u)
        out = np.zeros_like(u)
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        onp, nsh, gp = i64(pm.owned_nodes_ptr), i64(pm.n_shared), i64(pm.ghost_ptr)
        gi = np.ascontiguousarray(pm.ghost_idx, dtype=np.int32)
        P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        if two_pass:
            rp, ri = i64(pm.ghost_reduce_ptr), i64(pm.ghost_reduce_idx)
            rd = np.ascontiguousarray(pm.ghost_reduce_dest, dtype=np.int32)
            tail = (C.c_ssize_t(rd.size), P(rp, C.c_ssize_t), P(ri, C.c_ssize_t), P(rd, C.c_int32))
        else:
            tail = (C.c_ssize_t(0), None, None, None)
        self._call("packed_laplacian_apply_u16", C.c_int(element_type), C.c_ssize_t(pm.n_packs), C.c_ssize_t(pm.n_elements_per_pack),
                   C.c_ssize_t(pm.n_elements), C.c_ssize_t(pm.max_nodes_per_pack), _soa_ptrs(el, C.c_uint16), pf, P(onp, C.c_ssize_t),
                   P(nsh, C.c_ssize_t), P(gp, C.c_ssize_t), P(gi, C.c_int32), pu, out.ctypes.data_as(C.POINTER(self.crt)), *tail)
        return out

    def hex8_laplacian_diag(self, elements, points, nnodes):
        """hex8_laplacian_diag (operators/hex8/hex8_laplacian.hpp:36; geometry from the fp32 points, in scalar_t)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes, dtype=self.rt)
        self._call("hex8_laplacian_diag", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), out.ctypes.data_as(C.POINTER(self.crt)))
        return out

    def hex8_linear_elasticity_diag(self, elements, points, mu, lam, nnodes):
        """affine_hex8_linear_elasticity_diag (operators/hex8/hex8_linear_elasticity.cpp:1043); AoS [nnodes*3]."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 3, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("affine_hex8_linear_elasticity_diag", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32),
                   _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(3), op[0], op[1], op[2])
        return out

    def hex8_linear_elasticity_block_diag_sym(self, elements, points, mu, lam, nnodes):
        """affine_hex8_linear_elasticity_block_diag_sym (hex8_linear_elasticity.cpp:1100-1200); AoS [nnodes*6]."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 6, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(6)]
        self._call("affine_hex8_linear_elasticity_block_diag_sym", C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes),
                   _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(6), *op)
        return out

    def sshex8_linear_elasticity_diag(self, level, elements, points, mu, lam, nnodes):
        """affine_sshex8_linear_elasticity_diag (operators/sshex8/sshex8_linear_elasticity.cpp:971);
        `points` is SoA fp32 [3][nnodes] indexed by sshex8 node id (only the macro corners are read)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 3, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(3)]
        self._call("affine_sshex8_linear_elasticity_diag", C.c_int(level), C.c_ssize_t(el.shape[1]), C.c_ssize_t(nnodes),
                   _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam), C.c_ssize_t(3),
                   op[0], op[1], op[2])
        return out

    def sshex8_linear_elasticity_block_diag_sym(self, level, elements, points, mu, lam, nnodes):
        """affine_sshex8_linear_elasticity_block_diag_sym (:1102); AoS [nnodes*6] (00 01 02 11 12 22)."""
        el = np.ascontiguousarray(elements, dtype=np.int32)
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(nnodes * 6, dtype=self.rt)
        sz = out.itemsize
        op = [C.cast(out.ctypes.data + d * sz, C.POINTER(self.crt)) for d in range(6)]
        self._call("affine_sshex8_linear_elasticity_block_diag_sym", C.c_int(level), C.c_ssize_t(el.shape[1]),
                   C.c_ssize_t(nnodes), _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), self.crt(mu), self.crt(lam),
                   C.c_ssize_t(6), *op)
        return out

    # ---- multigrid transfer (operators/sshex8/sshex8_interpolate.cpp)
    def sshex8_element_node_incidence_count(self, level, stride, elements, nnodes):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        count = np.zeros(nnodes, dtype=np.uint16)
        self._call("sshex8_element_node_incidence_count", C.c_int(level), C.c_int(stride), C.c_ssize_t(el.shape[1]),
                   _soa_ptrs(el, C.c_int32), count.ctypes.data_as(C.POINTER(C.c_uint16)))
        return count

    def sshex8_hierarchical_restriction(self, level, elements, count, vec_size, frm, n_to):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        frm, pf = self._r(frm)
        to = np.zeros(n_to * vec_size, dtype=self.rt)
        self._call("sshex8_hierarchical_restriction", C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32),
                   count.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_int(vec_size), pf, to.ctypes.data_as(C.POINTER(self.crt)))
        return to

    def sshex8_hierarchical_prolongation(self, level, elements, vec_size, frm, n_to):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        frm, pf = self._r(frm)
        to = np.zeros(n_to * vec_size, dtype=self.rt)
        self._call("sshex8_hierarchical_prolongation", C.c_int(level), C.c_ssize_t(el.shape[1]), _soa_ptrs(el, C.c_int32),
                   C.c_int(vec_size), pf, to.ctypes.data_as(C.POINTER(self.crt)))
        return to

    def sshex8_restrict(self, from_level, from_stride, from_elements, count, to_level, to_stride, to_elements, vec_size,
                        frm, n_to):
        fe = np.ascontiguousarray(from_elements, dtype=np.int32)
        te = np.ascontiguousarray(to_elements, dtype=np.int32)
        frm, pf = self._r(frm)
        to = np.zeros(n_to * vec_size, dtype=self.rt)
        self._call("sshex8_restrict", C.c_ssize_t(fe.shape[1]), C.c_int(from_level), C.c_int(from_stride),
                   _soa_ptrs(fe, C.c_int32), count.ctypes.data_as(C.POINTER(C.c_uint16)), C.c_int(to_level),
                   C.c_int(to_stride), _soa_ptrs(te, C.c_int32), C.c_int(vec_size), pf,
                   to.ctypes.data_as(C.POINTER(self.crt)))
        return to

    def sshex8_prolongate(self, from_level, from_stride, from_elements, to_level, to_stride, to_elements, vec_size, frm,
                          n_to):
        fe = np.ascontiguousarray(from_elements, dtype=np.int32)
        te = np.ascontiguousarray(to_elements, dtype=np.int32)
        frm, pf = self._r(frm)
        to = np.zeros(n_to * vec_size, dtype=self.rt)
        self._call("sshex8_prolongate", C.c_ssize_t(fe.shape[1]), C.c_int(from_level), C.c_int(from_stride),
                   _soa_ptrs(fe, C.c_int32), C.c_int(to_level), C.c_int(to_stride), _soa_ptrs(te, C.c_int32),
                   C.c_int(vec_size), pf, to.ctypes.data_as(C.POINTER(self.crt)))
        return to

"""ctypes binding of the C ABI in include/b200fem.h (libb200fem.so, built in-tree by
`sfem_b200/csrc/Makefile`). This is plumbing for the tests and bench.py: the product is the shared
library and the C++ host layer in `sfem_b200/host/`. There is no CPU fallback -- importing this module
without the built library raises, and every call needs a CUDA device.

Vectors are torch CUDA tensors (torch = device memory + streams only); connectivity / geometry are
numpy host arrays exactly as the reference's operator-level C functions take them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200fem.so")

F32, F64 = 4, 8
TET4, HEX8, SSHEX8 = 4, 8, 800
TET10, MACRO_TET4 = 10, 1004
SCATTER_AUTO, SCATTER_ATOMIC, SCATTER_COLORED, SCATTER_PATCH = 0, 1, 2, 3
UNIQUE_ID_BYTES = 128
HALO_AUTO, HALO_PUSH, HALO_NCCL = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C sfem_b200/csrc`). b200fem has no CPU fallback."
    )
_lib = C.CDLL(LIB_PATH)

c_idx_pp = C.POINTER(C.POINTER(C.c_int32))
c_f32_pp = C.POINTER(C.POINTER(C.c_float))
vp, i32, f64, ssz = C.c_void_p, C.c_int, C.c_double, C.c_ssize_t

_SIG = {
    "b200fem_device_count": [C.POINTER(i32)],
    "b200fem_set_device": [i32],
    "b200fem_device_info": [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_size_t)],
    "b200fem_malloc": [C.POINTER(vp), C.c_size_t],
    "b200fem_free": [vp],
    "b200fem_host_malloc": [C.POINTER(vp), C.c_size_t],
    "b200fem_host_free": [vp],
    "b200fem_memcpy_h2d": [vp, vp, C.c_size_t, vp],
    "b200fem_memcpy_d2h": [vp, vp, C.c_size_t, vp],
    "b200fem_memcpy_d2d": [vp, vp, C.c_size_t, vp],
    "b200fem_stream_create": [C.POINTER(vp)],
    "b200fem_stream_destroy": [vp],
    "b200fem_stream_synchronize": [vp],
    "b200fem_device_synchronize": [],
    "b200fem_fff_fill": [i32, ssz, ssz, c_idx_pp, c_f32_pp, C.POINTER(C.c_float)],
    "b200fem_adjugate_and_det_fill": [i32, ssz, ssz, c_idx_pp, c_f32_pp, C.POINTER(C.c_float), C.POINTER(C.c_float)],
    "b200fem_laplacian_create": [C.POINTER(vp), i32, i32, ssz, ssz, c_idx_pp, C.POINTER(C.c_float), i32],
    "b200fem_linear_elasticity_create": [C.POINTER(vp), i32, i32, ssz, ssz, c_idx_pp, C.POINTER(C.c_float),
                                         C.POINTER(C.c_float), f64, f64, i32],
    "b200fem_laplacian_create_from_points": [C.POINTER(vp), i32, i32, ssz, ssz, c_idx_pp, c_f32_pp, i32],
    "b200fem_mass_create": [C.POINTER(vp), i32, ssz, ssz, c_idx_pp, c_f32_pp, i32],
    "b200fem_vector_laplacian_create": [C.POINTER(vp), i32, ssz, ssz, c_idx_pp, C.POINTER(C.c_float), i32, i32],
    "b200fem_sshex8_vector_laplacian_create": [C.POINTER(vp), i32, ssz, ssz, c_idx_pp, C.POINTER(C.c_float), i32, i32],
    "b200fem_sshex8_lumped_mass_create": [C.POINTER(vp), i32, ssz, ssz, c_idx_pp, c_f32_pp, i32],
    "b200fem_packed_mesh_create": [C.POINTER(vp), i32, ssz, ssz, c_idx_pp, ssz],
    "b200fem_packed_mesh_destroy": [vp],
    "b200fem_packed_mesh_sizes": [vp, C.POINTER(ssz), C.POINTER(ssz), C.POINTER(ssz), C.POINTER(ssz)],
    "b200fem_packed_mesh_arrays": [vp, C.POINTER(C.POINTER(C.POINTER(C.c_uint16))), C.POINTER(C.POINTER(ssz)), C.POINTER(C.POINTER(ssz)),
                                   C.POINTER(C.POINTER(ssz)), C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32))],
    "b200fem_packed_mesh_ghost_reduce": [vp, C.POINTER(ssz), C.POINTER(C.POINTER(ssz)), C.POINTER(C.POINTER(ssz)), C.POINTER(C.POINTER(C.c_int32))],
    "b200fem_packed_laplacian_set_two_pass": [vp, ssz, C.POINTER(ssz), C.POINTER(ssz), C.POINTER(C.c_int32)],
    "b200fem_packed_laplacian_create": [C.POINTER(vp), i32, ssz, ssz, ssz, ssz, ssz, C.POINTER(C.POINTER(C.c_uint16)), C.POINTER(C.c_float),
                                        C.POINTER(ssz), C.POINTER(ssz), C.POINTER(ssz), C.POINTER(C.c_int32), i32],
    "b200fem_linear_elasticity_create_from_points": [C.POINTER(vp), i32, i32, ssz, ssz, c_idx_pp, c_f32_pp, f64, f64, i32],
    "b200fem_sshex8_em_create": [C.POINTER(vp), i32, ssz, ssz, c_idx_pp, i32, vp, i32],
    "b200fem_sshex8_laplacian_element_matrix": [i32, ssz, C.POINTER(C.c_float), i32, vp],
    "b200fem_sshex8_linear_elasticity_element_matrix": [i32, ssz, C.POINTER(C.c_float), C.POINTER(C.c_float), f64, f64, i32, vp],
    "b200fem_op_destroy": [vp],
    "b200fem_op_apply": [vp, vp, vp, vp],
    "b200fem_op_apply_overwrite": [vp, vp, vp, vp],
    "b200fem_op_apply_range": [vp, vp, vp, ssz, ssz, i32, vp],
    "b200fem_op_apply_overwrite_after": [vp, vp, vp, ssz, vp, C.c_uint64, vp],
    "b200fem_op_diag": [vp, vp, vp],
    "b200fem_op_block_diag_sym": [vp, vp, vp],
    "b200fem_op_block_size": [vp],
    "b200fem_op_real_type": [vp],
    "b200fem_op_gpu_launches_per_apply": [vp, i32],
    "b200fem_op_set_option": [vp, C.c_char_p, i32],
    "b200fem_op_get_option": [vp, C.c_char_p, C.POINTER(i32)],
    "b200fem_constraint_nodes_copy_vec": [i32, ssz, vp, i32, i32, vp, vp, vp],
    "b200fem_constraint_nodes_to_value_vec": [i32, ssz, vp, i32, i32, f64, vp, vp],
    "b200fem_halo_pack": [i32, ssz, vp, i32, vp, vp, vp],
    "b200fem_halo_unpack_add": [i32, ssz, vp, i32, vp, vp, vp],
    "b200fem_blas_zeros": [i32, ssz, vp, vp],
    "b200fem_blas_values": [i32, ssz, f64, vp, vp],
    "b200fem_blas_copy": [i32, ssz, vp, vp, vp],
    "b200fem_blas_scal": [i32, ssz, f64, vp, vp],
    "b200fem_blas_axpy": [i32, ssz, f64, vp, vp, vp],
    "b200fem_blas_axpby": [i32, ssz, f64, vp, f64, vp, vp],
    "b200fem_blas_zaxpby": [i32, ssz, f64, vp, f64, vp, vp, vp],
    "b200fem_blas_xypaz": [i32, ssz, vp, vp, f64, vp, vp],
    "b200fem_blas_reciprocal": [i32, ssz, f64, vp, vp],
    "b200fem_blas_dot": [i32, ssz, vp, vp, C.POINTER(f64), vp],
    "b200fem_blas_norm2": [i32, ssz, vp, C.POINTER(f64), vp],
    "b200fem_blas_dot_async": [i32, ssz, vp, vp, vp, vp],
    "b200fem_cg_update_x_r_rtr": [i32, ssz, f64, vp, vp, vp, vp, C.POINTER(f64), vp],
    "b200fem_cg_update_x_r_rtr_async": [i32, ssz, f64, vp, vp, vp, vp, vp, vp],
    "b200fem_cg_update_p": [i32, ssz, f64, vp, vp, vp],
    "b200fem_cheb3_update": [i32, ssz, f64, f64, vp, vp, vp, vp, vp],
    "b200fem_sshex8_transfer_create": [C.POINTER(vp), ssz, i32, i32, c_idx_pp, ssz, i32, i32, c_idx_pp, ssz],
    "b200fem_transfer_destroy": [vp],
    "b200fem_transfer_incidence_count": [vp, C.POINTER(C.c_uint16)],
    "b200fem_transfer_set_incidence_count": [vp, C.POINTER(C.c_uint16)],
    "b200fem_transfer_restrict": [vp, i32, i32, vp, vp, vp],
    "b200fem_comm_unique_id": [vp],
    "b200fem_comm_create": [C.POINTER(vp), i32, i32, vp],
    "b200fem_comm_destroy": [vp],
    "b200fem_comm_rank": [vp],
    "b200fem_comm_size": [vp],
    "b200fem_comm_error": [vp],
    "b200fem_comm_allreduce_sum": [vp, vp, i32, vp],
    "b200fem_comm_sum_to_host": [vp, vp, i32, C.POINTER(f64), vp],
    "b200fem_comm_allreduce_host": [vp, C.POINTER(f64), i32, i32, vp],
    "b200fem_comm_barrier": [vp, vp],
    "b200fem_halo_create": [C.POINTER(vp), vp, i32, i32, ssz, ssz, i32, vp, vp, vp, i32, vp, vp, vp, i32],
    "b200fem_halo_destroy": [vp],
    "b200fem_halo_gather": [vp, vp, vp],
    "b200fem_halo_set_recv_index": [vp, vp, vp],
    "b200fem_halo_gather_begin": [vp, vp, vp],
    "b200fem_halo_gather_end": [vp, vp, vp],
    "b200fem_halo_transport": [vp],
    "b200fem_halo_launches_per_gather": [vp],
    "b200fem_transfer_prolongate": [vp, i32, i32, vp, vp, vp],
    "b200fem_block_sym_jacobi_setup": [i32, ssz, vp, vp, vp, f64, vp, vp],
    "b200fem_block_jacobi_apply": [i32, ssz, vp, vp, vp, vp],
    "b200fem_block_sym_jacobi_setup_mixed": [ssz, vp, vp, vp, f64, vp, vp],
    "b200fem_block_jacobi_apply_mixed": [ssz, vp, vp, vp, vp],
}
for _name, _args in _SIG.items():
    _fn = getattr(_lib, _name)
    _fn.argtypes = _args
    _fn.restype = i32
_lib.b200fem_last_error.restype = C.c_char_p
_lib.b200fem_op_rows.argtypes = [vp]
_lib.b200fem_op_rows.restype = ssz
_lib.b200fem_op_algorithmic_bytes.argtypes = [vp]
_lib.b200fem_op_algorithmic_bytes.restype = f64
for _name in ("b200fem_transfer_fine_nodes", "b200fem_transfer_coarse_nodes", "b200fem_halo_values_sent", "b200fem_op_nelements"):
    getattr(_lib, _name).argtypes = [vp]
    getattr(_lib, _name).restype = ssz
_lib.b200fem_op_streamed_bytes.argtypes = [vp, i32]
_lib.b200fem_op_streamed_bytes.restype = f64


class B200FemError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        raise B200FemError(f"{what}: {_lib.b200fem_last_error().decode()}")


def raw():
    """The ctypes CDLL (for symbol checks)."""
    return _lib


def declared_symbols():
    """Every entry point include/b200fem.h declares (parsed from the header)."""
    import re

    hdr = open(os.path.join(HERE, "..", "include", "b200fem.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200fem_[a-z0-9_]+)\s*\(", hdr)))


def _soa_ptrs(arr: np.ndarray, ctype):
    k = arr.shape[0]
    ptrs = (C.POINTER(ctype) * k)()
    for i in range(k):
        ptrs[i] = arr[i].ctypes.data_as(C.POINTER(ctype))
    return ptrs


def _rt(dtype) -> int:
    import torch

    if dtype in (torch.float64, np.float64, "f64", F64):
        return F64
    if dtype in (torch.float32, np.float32, "f32", F32):
        return F32
    raise ValueError(f"unsupported real type {dtype}")


def _stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _vec(t, n=None):
    assert t.is_cuda and t.is_contiguous(), "vectors must be contiguous CUDA tensors"
    if n is not None:
        assert t.numel() >= n
    return t


# --------------------------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------------------------
def fff_fill(element_type: int, elements: np.ndarray, points: np.ndarray) -> np.ndarray:
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros(el.shape[1] * 6, dtype=np.float32)
    check(_lib.b200fem_fff_fill(element_type, el.shape[1], pts.shape[1], _soa_ptrs(el, C.c_int32),
                                _soa_ptrs(pts, C.c_float), out.ctypes.data_as(C.POINTER(C.c_float))), "fff_fill")
    return out


def adjugate_and_det_fill(element_type: int, elements: np.ndarray, points: np.ndarray):
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    adj = np.zeros(el.shape[1] * 9, dtype=np.float32)
    det = np.zeros(el.shape[1], dtype=np.float32)
    check(_lib.b200fem_adjugate_and_det_fill(element_type, el.shape[1], pts.shape[1], _soa_ptrs(el, C.c_int32),
                                             _soa_ptrs(pts, C.c_float), adj.ctypes.data_as(C.POINTER(C.c_float)),
                                             det.ctypes.data_as(C.POINTER(C.c_float))), "adjugate_and_det_fill")
    return adj, det


# --------------------------------------------------------------------------------------------
# operator plans
# --------------------------------------------------------------------------------------------
class Op:
    """Owner of a b200fem_op_t. `apply(x, y)` accumulates (y += A x) like sfem::Op::apply."""

    def __init__(self, handle, nnodes):
        self._h = handle
        self.nnodes = nnodes

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:  # module globals are gone at interpreter shutdown
            _lib.b200fem_op_destroy(h)

    @property
    def handle(self):
        return self._h

    @property
    def rows(self):
        return int(_lib.b200fem_op_rows(self._h))

    @property
    def block_size(self):
        return int(_lib.b200fem_op_block_size(self._h))

    @property
    def algorithmic_bytes(self):
        return float(_lib.b200fem_op_algorithmic_bytes(self._h))

    def streamed_bytes(self, overwrite=False):
        return float(_lib.b200fem_op_streamed_bytes(self._h, int(overwrite)))

    def launches(self, overwrite=False):
        return int(_lib.b200fem_op_gpu_launches_per_apply(self._h, int(overwrite)))

    def set_option(self, key: str, value: int):
        check(_lib.b200fem_op_set_option(self._h, key.encode(), int(value)), f"set_option({key})")

    def get_option(self, key: str) -> int:
        v = i32(0)
        check(_lib.b200fem_op_get_option(self._h, key.encode(), C.byref(v)), f"get_option({key})")
        return v.value

    def apply(self, x, y):
        _vec(x, self.rows), _vec(y, self.rows)
        check(_lib.b200fem_op_apply(self._h, _p(x), _p(y), _stream()), "op_apply")
        return y

    def apply_overwrite(self, x, y):
        _vec(x, self.rows), _vec(y, self.rows)
        check(_lib.b200fem_op_apply_overwrite(self._h, _p(x), _p(y), _stream()), "op_apply_overwrite")
        return y

    def apply_range(self, x, y, e_begin: int, e_end: int, mode: int = 0):
        """Elements [e_begin, e_end) only. mode 0: accumulate; 1 / 2: first part / continuation of an overwriting apply."""
        _vec(x, self.rows), _vec(y, self.rows)
        check(_lib.b200fem_op_apply_range(self._h, _p(x), _p(y), e_begin, e_end, mode, _stream()), "op_apply_range")
        return y

    def apply_overwrite_after(self, x, y, e_dependent: int, ready_flag, ready_value: int):
        """y = A x in one launch whose elements >= e_dependent wait inside the kernel until the device word `ready_flag`
        (a uint64 / int64 CUDA tensor of one element, or None) has reached ready_value."""
        _vec(x, self.rows), _vec(y, self.rows)
        check(_lib.b200fem_op_apply_overwrite_after(self._h, _p(x), _p(y), e_dependent, _p(ready_flag) if ready_flag is not None else None,
                                                    ready_value, _stream()), "op_apply_overwrite_after")
        return y

    def diag(self, d):
        _vec(d, self.rows)
        check(_lib.b200fem_op_diag(self._h, _p(d), _stream()), "op_diag")
        return d

    def block_diag_sym(self, d):
        _vec(d, self.nnodes * 6)
        check(_lib.b200fem_op_block_diag_sym(self._h, _p(d), _stream()), "op_block_diag_sym")
        return d


def laplacian(element_type: int, elements: np.ndarray, fff: np.ndarray, nnodes: int, dtype, level: int = 0) -> Op:
    el = np.ascontiguousarray(elements, dtype=np.int32)
    f = np.ascontiguousarray(fff, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_laplacian_create(C.byref(h), element_type, level, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32),
                                        f.ctypes.data_as(C.POINTER(C.c_float)), _rt(dtype)), "laplacian_create")
    return Op(h, nnodes)


def linear_elasticity(element_type: int, elements: np.ndarray, adjugate: np.ndarray, determinant: np.ndarray, mu: float,
                      lam: float, nnodes: int, dtype, level: int = 0) -> Op:
    el = np.ascontiguousarray(elements, dtype=np.int32)
    a = np.ascontiguousarray(adjugate, dtype=np.float32)
    d = np.ascontiguousarray(determinant, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_linear_elasticity_create(C.byref(h), element_type, level, el.shape[1], nnodes,
                                                _soa_ptrs(el, C.c_int32), a.ctypes.data_as(C.POINTER(C.c_float)),
                                                d.ctypes.data_as(C.POINTER(C.c_float)), mu, lam, _rt(dtype)),
          "linear_elasticity_create")
    return Op(h, nnodes)


def _np_rt(dtype):
    return np.float64 if _rt(dtype) == F64 else np.float32


def laplacian_from_points(element_type: int, elements: np.ndarray, points: np.ndarray, nnodes: int, dtype, level: int = 0) -> Op:
    """TET4: tet4_laplacian_apply (FFF from the points per element); SSHEX8: sshex8_laplacian_apply (non-affine, q6)."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_laplacian_create_from_points(C.byref(h), element_type, level, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32),
                                                    _soa_ptrs(pts, C.c_float), _rt(dtype)), "laplacian_create_from_points")
    return Op(h, nnodes)


class PackedMesh:
    """smesh::PackedMesh<uint16_t> as PackedLaplacian::apply reads it (frontend/ops/sfem_PackedLaplacian.cpp:875-930), built by
    b200fem_packed_mesh_create. numpy copies of the arrays; nodes are renumbered: packed node i is input node new_to_old[i]."""

    def __init__(self, elements: np.ndarray, nnodes: int, n_elements_per_pack: int = 0):
        el = np.ascontiguousarray(elements, dtype=np.int32)
        nxe, ne = el.shape
        h = vp()
        check(_lib.b200fem_packed_mesh_create(C.byref(h), nxe, ne, nnodes, _soa_ptrs(el, C.c_int32), n_elements_per_pack), "packed_mesh_create")
        try:
            a, b, c, d = ssz(), ssz(), ssz(), ssz()
            check(_lib.b200fem_packed_mesh_sizes(h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "packed_mesh_sizes")
            self.n_packs, self.n_elements_per_pack, self.max_nodes_per_pack, self.n_ghost_entries = a.value, b.value, c.value, d.value
            rows = C.POINTER(C.POINTER(C.c_uint16))()
            onp, nsh, gp = C.POINTER(ssz)(), C.POINTER(ssz)(), C.POINTER(ssz)()
            gi, n2o = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
            check(_lib.b200fem_packed_mesh_arrays(h, C.byref(rows), C.byref(onp), C.byref(nsh), C.byref(gp), C.byref(gi), C.byref(n2o)),
                  "packed_mesh_arrays")
            cp = lambda ptr, n, dt: np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dtype=dt)
            self.elements = np.stack([cp(rows[v], ne, np.uint16) for v in range(nxe)]) if ne else np.zeros((nxe, 0), dtype=np.uint16)
            self.owned_nodes_ptr = cp(onp, self.n_packs + 1, np.int64)
            self.n_shared = cp(nsh, self.n_packs, np.int64)
            self.ghost_ptr = cp(gp, self.n_packs + 1, np.int64)
            self.ghost_idx = cp(gi, self.n_ghost_entries, np.int32)
            self.new_to_old = cp(n2o, nnodes, np.int32)
            nr, rp, ri, rd = ssz(), C.POINTER(ssz)(), C.POINTER(ssz)(), C.POINTER(C.c_int32)()
            check(_lib.b200fem_packed_mesh_ghost_reduce(h, C.byref(nr), C.byref(rp), C.byref(ri), C.byref(rd)), "packed_mesh_ghost_reduce")
            self.ghost_reduce_ptr = cp(rp, nr.value + 1, np.int64) if nr.value else np.zeros(1, dtype=np.int64)
            self.ghost_reduce_idx = cp(ri, self.n_ghost_entries, np.int64)
            self.ghost_reduce_dest = cp(rd, nr.value, np.int32)
        finally:
            _lib.b200fem_packed_mesh_destroy(h)
        self.n_nodes, self.n_elements = nnodes, ne

    def global_elements(self) -> np.ndarray:
        """The packed connectivity expanded to (packed) global node ids [nxe][ne] -- what the local indices stand for."""
        ne = self.n_elements
        pack = np.arange(ne) // max(self.n_elements_per_pack, 1)
        base = self.owned_nodes_ptr[pack]
        n_owned = self.owned_nodes_ptr[pack + 1] - base
        out = np.empty(self.elements.shape, dtype=np.int32)
        for v in range(self.elements.shape[0]):
            l = self.elements[v].astype(np.int64)
            own = l < n_owned
            g = np.where(own, base + l, 0)
            gh = ~own
            g[gh] = self.ghost_idx[self.ghost_ptr[pack[gh]] + l[gh] - n_owned[gh]]
            out[v] = g
        return out


def packed_laplacian(element_type: int, pm: "PackedMesh", fff: np.ndarray, dtype, two_pass: bool = False) -> Op:
    """PackedLaplacian on a PackedMesh (TET4 / HEX8 / TET10 / MACRO_TET4); vectors in the packed node numbering.
    two_pass: hand over the ghost-reduce lists and select the atomics-free two-pass apply (SFEM_PACKED_TWO_PASS)."""
    el = np.ascontiguousarray(pm.elements, dtype=np.uint16)
    f = np.ascontiguousarray(fff, dtype=np.float32)
    onp = np.ascontiguousarray(pm.owned_nodes_ptr, dtype=np.int64)
    nsh = np.ascontiguousarray(pm.n_shared, dtype=np.int64)
    gp = np.ascontiguousarray(pm.ghost_ptr, dtype=np.int64)
    gi = np.ascontiguousarray(pm.ghost_idx, dtype=np.int32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    h = vp()
    check(_lib.b200fem_packed_laplacian_create(C.byref(h), element_type, pm.n_packs, pm.n_elements_per_pack, pm.n_elements, pm.n_nodes,
                                               pm.max_nodes_per_pack, _soa_ptrs(el, C.c_uint16), P(f, C.c_float), P(onp, ssz), P(nsh, ssz),
                                               P(gp, ssz), P(gi, C.c_int32), _rt(dtype)), "packed_laplacian_create")
    op = Op(h, pm.n_nodes)
    if two_pass:
        rp = np.ascontiguousarray(pm.ghost_reduce_ptr, dtype=np.int64)
        ri = np.ascontiguousarray(pm.ghost_reduce_idx, dtype=np.int64)
        rd = np.ascontiguousarray(pm.ghost_reduce_dest, dtype=np.int32)
        check(_lib.b200fem_packed_laplacian_set_two_pass(h, rd.size, P(rp, ssz), P(ri, ssz), P(rd, C.c_int32)), "packed_laplacian_set_two_pass")
        op.set_option("packed_two_pass", 1)
    return op


def vector_laplacian(element_type: int, elements: np.ndarray, fff: np.ndarray, nnodes: int, vector_size: int, dtype) -> Op:
    """HEX8: affine_hex8_vector_laplacian_apply_fff, the scalar Laplacian on every component of an AoS field."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    f = np.ascontiguousarray(fff, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_vector_laplacian_create(C.byref(h), element_type, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32),
                                               f.ctypes.data_as(C.POINTER(C.c_float)), vector_size, _rt(dtype)), "vector_laplacian_create")
    return Op(h, nnodes)


def sshex8_vector_laplacian(level: int, elements: np.ndarray, fff: np.ndarray, nnodes: int, vector_size: int, dtype) -> Op:
    """affine_sshex8_vector_laplacian_apply_fff: the scalar sshex8 Laplacian on every component of an AoS field."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    f = np.ascontiguousarray(fff, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_sshex8_vector_laplacian_create(C.byref(h), level, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32),
                                                      f.ctypes.data_as(C.POINTER(C.c_float)), vector_size, _rt(dtype)),
          "sshex8_vector_laplacian_create")
    return Op(h, nnodes)


def sshex8_lumped_mass(level: int, elements: np.ndarray, points: np.ndarray, nnodes: int, dtype) -> Op:
    """affine_sshex8_mass_lumped: diag only."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_sshex8_lumped_mass_create(C.byref(h), level, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32),
                                                 _soa_ptrs(pts, C.c_float), _rt(dtype)), "sshex8_lumped_mass_create")
    return Op(h, nnodes)


def mass(element_type: int, elements: np.ndarray, points: np.ndarray, nnodes: int, dtype) -> Op:
    """Mass / LumpedMass: apply = hex8_apply_mass (HEX8 only, like the reference), diag = the lumped mass (HEX8, TET4)."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_mass_create(C.byref(h), element_type, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float),
                                   _rt(dtype)), "mass_create")
    return Op(h, nnodes)


def linear_elasticity_from_points(element_type: int, elements: np.ndarray, points: np.ndarray, mu: float, lam: float, nnodes: int,
                                  dtype, level: int = 0) -> Op:
    """HEX8: hex8_linear_elasticity_apply, the isoparametric variant (Jacobian per Gauss point)."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    h = vp()
    check(_lib.b200fem_linear_elasticity_create_from_points(C.byref(h), element_type, level, el.shape[1], nnodes,
                                                            _soa_ptrs(el, C.c_int32), _soa_ptrs(pts, C.c_float), mu, lam, _rt(dtype)),
          "linear_elasticity_create_from_points")
    return Op(h, nnodes)


def sshex8_laplacian_element_matrix(level: int, fff: np.ndarray, dtype) -> np.ndarray:
    """[ne, 64] element matrices of ss:em:Laplacian from the macro FFF (sshex8_laplacian_element_matrix)."""
    f = np.ascontiguousarray(fff, dtype=np.float32).reshape(-1, 6)
    out = np.zeros((f.shape[0], 64), dtype=_np_rt(dtype))
    check(_lib.b200fem_sshex8_laplacian_element_matrix(level, f.shape[0], f.ctypes.data_as(C.POINTER(C.c_float)), _rt(dtype),
                                                       out.ctypes.data), "sshex8_laplacian_element_matrix")
    return out


def sshex8_linear_elasticity_element_matrix(level: int, adjugate: np.ndarray, determinant: np.ndarray, mu: float, lam: float,
                                            dtype) -> np.ndarray:
    """[ne, 576] element matrices of ss:em:LinearElasticity (sshex8_linear_elasticity_element_matrix)."""
    a = np.ascontiguousarray(adjugate, dtype=np.float32).reshape(-1, 9)
    d = np.ascontiguousarray(determinant, dtype=np.float32)
    out = np.zeros((a.shape[0], 576), dtype=_np_rt(dtype))
    check(_lib.b200fem_sshex8_linear_elasticity_element_matrix(level, a.shape[0], a.ctypes.data_as(C.POINTER(C.c_float)),
                                                               d.ctypes.data_as(C.POINTER(C.c_float)), mu, lam, _rt(dtype),
                                                               out.ctypes.data), "sshex8_linear_elasticity_element_matrix")
    return out


def sshex8_em(level: int, elements: np.ndarray, element_matrix: np.ndarray, block_size: int, nnodes: int, dtype) -> Op:
    """Stored-element-matrix operator (ss:em:Laplacian / ss:em:LinearElasticity): one (8 bs)^2 matrix per macro element."""
    el = np.ascontiguousarray(elements, dtype=np.int32)
    m = np.ascontiguousarray(element_matrix, dtype=_np_rt(dtype))
    assert m.size == el.shape[1] * (8 * block_size) ** 2
    h = vp()
    check(_lib.b200fem_sshex8_em_create(C.byref(h), level, el.shape[1], nnodes, _soa_ptrs(el, C.c_int32), block_size, m.ctypes.data,
                                        _rt(dtype)), "sshex8_em_create")
    return Op(h, nnodes)


# --------------------------------------------------------------------------------------------
# vector kernels (torch tensors in, same semantics as sfem::BLAS<T>)
# --------------------------------------------------------------------------------------------
class blas:
    @staticmethod
    def zeros(x):
        check(_lib.b200fem_blas_zeros(_rt(x.dtype), x.numel(), _p(x), _stream()), "zeros")

    @staticmethod
    def values(v, x):
        check(_lib.b200fem_blas_values(_rt(x.dtype), x.numel(), v, _p(x), _stream()), "values")

    @staticmethod
    def copy(src, dst):
        check(_lib.b200fem_blas_copy(_rt(src.dtype), src.numel(), _p(src), _p(dst), _stream()), "copy")

    @staticmethod
    def scal(a, x):
        check(_lib.b200fem_blas_scal(_rt(x.dtype), x.numel(), a, _p(x), _stream()), "scal")

    @staticmethod
    def axpy(a, x, y):
        check(_lib.b200fem_blas_axpy(_rt(x.dtype), x.numel(), a, _p(x), _p(y), _stream()), "axpy")

    @staticmethod
    def axpby(a, x, b, y):
        check(_lib.b200fem_blas_axpby(_rt(x.dtype), x.numel(), a, _p(x), b, _p(y), _stream()), "axpby")

    @staticmethod
    def zaxpby(a, x, b, y, z):
        check(_lib.b200fem_blas_zaxpby(_rt(x.dtype), x.numel(), a, _p(x), b, _p(y), _p(z), _stream()), "zaxpby")

    @staticmethod
    def xypaz(x, y, a, z):
        check(_lib.b200fem_blas_xypaz(_rt(x.dtype), x.numel(), _p(x), _p(y), a, _p(z), _stream()), "xypaz")

    @staticmethod
    def reciprocal(a, x):
        check(_lib.b200fem_blas_reciprocal(_rt(x.dtype), x.numel(), a, _p(x), _stream()), "reciprocal")

    @staticmethod
    def dot(x, y) -> float:
        r = f64(0)
        check(_lib.b200fem_blas_dot(_rt(x.dtype), x.numel(), _p(x), _p(y), C.byref(r), _stream()), "dot")
        return r.value

    @staticmethod
    def norm2(x) -> float:
        r = f64(0)
        check(_lib.b200fem_blas_norm2(_rt(x.dtype), x.numel(), _p(x), C.byref(r), _stream()), "norm2")
        return r.value

    @staticmethod
    def dot_async(x, y, result_dev):
        check(_lib.b200fem_blas_dot_async(_rt(x.dtype), x.numel(), _p(x), _p(y), _p(result_dev), _stream()), "dot_async")

    @staticmethod
    def cg_update_x_r_rtr(alpha, p, Ap, x, r) -> float:
        out = f64(0)
        check(_lib.b200fem_cg_update_x_r_rtr(_rt(x.dtype), x.numel(), alpha, _p(p), _p(Ap), _p(x), _p(r), C.byref(out),
                                              _stream()), "cg_update_x_r_rtr")
        return out.value

    @staticmethod
    def cg_update_p(beta, z, p):
        check(_lib.b200fem_cg_update_p(_rt(p.dtype), p.numel(), beta, _p(z), _p(p), _stream()), "cg_update_p")

    @staticmethod
    def cheb3_update(alpha, beta, rhs, Ax, p, x):
        check(_lib.b200fem_cheb3_update(_rt(x.dtype), x.numel(), alpha, beta, _p(rhs), _p(Ax), _p(p), _p(x), _stream()),
              "cheb3_update")


def constraint_nodes_copy_vec(idx, block_size, component, source, dest):
    check(_lib.b200fem_constraint_nodes_copy_vec(_rt(source.dtype), idx.numel(), _p(idx), block_size, component,
                                                 _p(source), _p(dest), _stream()), "constraint_nodes_copy_vec")


def constraint_nodes_to_value_vec(idx, block_size, component, value, values):
    check(_lib.b200fem_constraint_nodes_to_value_vec(_rt(values.dtype), idx.numel(), _p(idx), block_size, component,
                                                     value, _p(values), _stream()), "constraint_nodes_to_value_vec")


def halo_pack(idx, block_size, src, dst):
    """dst[i*bs + c] = src[idx[i]*bs + c] (owner side of smesh::Exchange::gather)."""
    check(_lib.b200fem_halo_pack(_rt(src.dtype), idx.numel(), _p(idx), block_size, _p(src), _p(dst), _stream()), "halo_pack")


def halo_unpack_add(idx, block_size, src, dst):
    """dst[idx[i]*bs + c] += src[i*bs + c] (owner side of smesh::Exchange::scatter_add)."""
    check(_lib.b200fem_halo_unpack_add(_rt(src.dtype), idx.numel(), _p(idx), block_size, _p(src), _p(dst), _stream()),
          "halo_unpack_add")


# --------------------------------------------------------------------------------------------
# multigrid transfer + smoothers
# --------------------------------------------------------------------------------------------
class Transfer:
    """Owner of a b200fem_transfer_t: restriction / prolongation between two lattices of one sshex8 macro
    mesh (sshex8_restrict / sshex8_prolongate / the hierarchical pair of the reference)."""

    def __init__(self, fine_level, fine_stride, fine_elements, n_fine_nodes, coarse_level, coarse_stride,
                 coarse_elements, n_coarse_nodes):
        fe = np.ascontiguousarray(fine_elements, dtype=np.int32)
        ce = np.ascontiguousarray(coarse_elements, dtype=np.int32)
        h = vp()
        check(_lib.b200fem_sshex8_transfer_create(C.byref(h), fe.shape[1], fine_level, fine_stride, _soa_ptrs(fe, C.c_int32),
                                                  n_fine_nodes, coarse_level, coarse_stride, _soa_ptrs(ce, C.c_int32),
                                                  n_coarse_nodes), "sshex8_transfer_create")
        self._h = h
        self.n_fine, self.n_coarse = n_fine_nodes, n_coarse_nodes

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.b200fem_transfer_destroy(h)

    @property
    def handle(self):
        return self._h

    def incidence_count(self) -> np.ndarray:
        out = np.zeros(self.n_fine, dtype=np.uint16)
        check(_lib.b200fem_transfer_incidence_count(self._h, out.ctypes.data_as(C.POINTER(C.c_uint16))), "incidence_count")
        return out

    def set_incidence_count(self, count: np.ndarray):
        """Distributed meshes: hand the globally completed table back (ghost entries from their owners)."""
        c = np.ascontiguousarray(count, dtype=np.uint16)
        assert c.size == self.n_fine
        check(_lib.b200fem_transfer_set_incidence_count(self._h, c.ctypes.data_as(C.POINTER(C.c_uint16))), "set_incidence_count")

    def restrict(self, frm, to, vec_size=1):
        """to[coarse] += R from[fine] (the caller zeroes `to`, like the reference)."""
        _vec(frm, self.n_fine * vec_size), _vec(to, self.n_coarse * vec_size)
        check(_lib.b200fem_transfer_restrict(self._h, _rt(frm.dtype), vec_size, _p(frm), _p(to), _stream()), "restrict")
        return to

    def prolongate(self, frm, to, vec_size=1):
        """to[fine] = P from[coarse]."""
        _vec(frm, self.n_coarse * vec_size), _vec(to, self.n_fine * vec_size)
        check(_lib.b200fem_transfer_prolongate(self._h, _rt(frm.dtype), vec_size, _p(frm), _p(to), _stream()), "prolongate")
        return to


def block_sym_jacobi_setup(sym_diag6, relaxation, inv_diag9, shift6=None, mask=None):
    n_blocks = sym_diag6.numel() // 6
    check(_lib.b200fem_block_sym_jacobi_setup(_rt(sym_diag6.dtype), n_blocks, _p(sym_diag6),
                                              _p(shift6) if shift6 is not None else None,
                                              _p(mask) if mask is not None else None, relaxation, _p(inv_diag9), _stream()),
          "block_sym_jacobi_setup")
    return inv_diag9


def block_jacobi_apply(inv_diag9, x, y):
    check(_lib.b200fem_block_jacobi_apply(_rt(x.dtype), x.numel() // 3, _p(inv_diag9), _p(x), _p(y), _stream()),
          "block_jacobi_apply")
    return y

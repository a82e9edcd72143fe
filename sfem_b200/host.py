"""ctypes front of the C++ host layer (sfem_b200/host/b200fem_host.hpp through host_capi.cpp ->
libb200fem_host.so): Function, ConjugateGradient, Chebyshev3, StationaryIteration, ShiftableJacobi,
ShiftableBlockSymJacobi and Multigrid with the reference's interfaces (algebra/sfem_cg.hpp, sfem_Chebyshev3.hpp,
sfem_Stationary.hpp, sfem_ShiftableJacobi.hpp, sfem_Multigrid.hpp, frontend/sfem_Function.cpp), running entirely on
the CUDA kernels behind include/b200fem.h. Plumbing for tests and bench.py; no CPU fallback.

The object-graph calls mirror oracle/ref_solvers.py (the reference's own solver headers) one to one, so a parity
test builds the same solver twice and compares iteration counts and solutions.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib as _l

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200fem_host.so")
if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `make -C sfem_b200/host` (after sfem_b200/csrc).")
_h = C.CDLL(LIB_PATH)

vp, i32, f64, ssz = C.c_void_p, C.c_int, C.c_double, C.c_ssize_t
APPLY_CB = C.CFUNCTYPE(i32, vp, vp, vp)
REDUCE_CB = C.CFUNCTYPE(f64, vp, f64)

_SIG = {
    "b200host_blas": (vp, [i32, REDUCE_CB, vp]),
    "b200host_blas_launches": (C.c_long, [vp]),
    "b200host_blas_set_allocator": (i32, [vp, C.CFUNCTYPE(vp, vp, C.c_size_t), C.CFUNCTYPE(None, vp, vp), vp]),
    "b200host_function": (vp, [i32, ssz, i32]),
    "b200host_function_add_operator": (i32, [vp, vp]),
    "b200host_function_add_dirichlet": (i32, [vp, vp, ssz, i32, f64]),
    "b200host_function_apply_constraints": (i32, [vp, vp]),
    "b200host_function_apply_zero_constraints": (i32, [vp, vp]),
    "b200host_function_set_value_to_constrained_dofs": (i32, [vp, f64, vp]),
    "b200host_function_hessian_diag": (i32, [vp, vp]),
    "b200host_function_hessian_block_diag_sym": (i32, [vp, vp]),
    "b200host_function_launches_per_apply": (i32, [vp, i32]),
    "b200host_callback_op": (vp, [i32, ssz, ssz, APPLY_CB, APPLY_CB, vp]),
    "b200host_restriction": (vp, [i32, vp, i32, vp]),
    "b200host_prolongation": (vp, [i32, vp, i32, vp]),
    "b200host_cg": (vp, [vp, vp, vp, f64, f64, i32]),
    "b200host_jacobi": (vp, [vp, vp, ssz, f64]),
    "b200host_block_sym_jacobi": (vp, [vp, vp, ssz, vp, f64]),
    "b200host_mixed_block_sym_jacobi": (vp, [vp, ssz, vp, f64]),
    "b200host_stationary": (vp, [vp, vp, vp, i32]),
    "b200host_cheb3": (vp, [vp, vp, f64, i32, f64, f64, i32, i32]),
    "b200host_cheb3_init_with_ones": (i32, [vp]),
    "b200host_cheb3_init": (i32, [vp, vp]),
    "b200host_cheb3_eig_max": (f64, [vp]),
    "b200host_cheb3_power_method_iterations": (i32, [vp]),
    "b200host_cheb3_set_initial_guess_zero": (i32, [vp, i32]),
    "b200host_mg": (vp, [vp, i32, f64, i32, i32]),
    "b200host_mg_add_level": (i32, [vp, vp, vp, vp, vp]),
    "b200host_comm_unique_id": (i32, [vp]),
    "b200host_comm": (vp, [i32, i32, vp]),
    "b200host_comm_rank": (i32, [vp]),
    "b200host_comm_size": (i32, [vp]),
    "b200host_comm_sum": (f64, [vp, f64]),
    "b200host_comm_max": (f64, [vp, f64]),
    "b200host_comm_error": (i32, [vp]),
    "b200host_comm_barrier": (None, [vp]),
    "b200host_halo": (vp, [vp, i32, i32, ssz, ssz, i32, vp, vp, vp, i32, vp, vp, vp, i32]),
    "b200host_halo_gather": (i32, [vp, vp]),
    "b200host_halo_set_recv_index": (i32, [vp, vp, vp]),
    "b200host_halo_transport": (i32, [vp]),
    "b200host_halo_values_sent": (ssz, [vp]),
    "b200host_halo_launches_per_gather": (i32, [vp]),
    "b200host_parallel_op": (vp, [vp, vp, ssz, ssz]),
    "b200host_parallel_op_set_independent_range": (i32, [vp, ssz, ssz]),
    "b200host_blas_set_comm": (i32, [vp, vp]),
    "b200host_blas_reserve_allocation": (i32, [vp, ssz, ssz]),
    "b200host_blas_reductions": (C.c_long, [vp]),
    "b200host_blas_dot": (f64, [vp, ssz, vp, vp]),
    "b200host_mesh_last_error": (C.c_char_p, []),
    "b200host_mesh_read": (vp, [C.c_char_p]),
    "b200host_mesh_destroy": (None, [vp]),
    "b200host_mesh_n_elements": (ssz, [vp]),
    "b200host_mesh_n_nodes": (ssz, [vp]),
    "b200host_mesh_nxe": (i32, [vp]),
    "b200host_mesh_dim": (i32, [vp]),
    "b200host_mesh_element_type": (C.c_char_p, [vp]),
    "b200host_mesh_elements": (C.POINTER(C.c_int32), [vp, i32]),
    "b200host_mesh_points": (C.POINTER(C.c_float), [vp, i32]),
    "b200host_mesh_write": (i32, [C.c_char_p, C.c_char_p, i32, ssz, ssz, i32, C.POINTER(C.POINTER(C.c_int32)),
                                  C.POINTER(C.POINTER(C.c_float))]),
    "b200host_apply": (i32, [vp, vp, vp]),
    "b200host_apply_overwrite": (i32, [vp, vp, vp]),
    "b200host_rows": (ssz, [vp]),
    "b200host_iterations": (i32, [vp]),
    "b200host_residual_history": (i32, [vp, C.POINTER(f64), i32]),
    "b200host_destroy": (None, [vp]),
}
for _n, (_r, _a) in _SIG.items():
    getattr(_h, _n).restype = _r
    getattr(_h, _n).argtypes = _a


def raw():
    return _h


def exported_symbols():
    return sorted(_SIG)


class Handle:
    def __init__(self, ptr, keep=()):
        if not ptr:
            raise _l.B200FemError("host layer returned NULL: " + _l.raw().b200fem_last_error().decode())
        self.ptr, self.keep = ptr, list(keep)

    def __del__(self):
        try:
            if self.ptr:
                _h.b200host_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Blas(Handle):
    """DeviceBLAS<T>; `allreduce(value) -> global sum` makes every dot / norm a distributed one."""

    def __init__(self, dtype, allreduce=None):
        self.rt = _l._rt(dtype)
        cb = REDUCE_CB(lambda _ctx, v: float(allreduce(v))) if allreduce else C.cast(None, REDUCE_CB)
        super().__init__(_h.b200host_blas(self.rt, cb, None), keep=[cb, allreduce])

    def launches(self) -> int:
        return _h.b200host_blas_launches(self.ptr)

    def reductions(self) -> int:
        return _h.b200host_blas_reductions(self.ptr)

    def set_comm(self, comm: "Communicator"):
        """dots / norms / fused CG reductions become device-side sums over the communicator's ranks"""
        self.keep.append(comm)
        _l.check(_h.b200host_blas_set_comm(self.ptr, comm.ptr), "blas_set_comm")

    def reserve_allocation(self, n_owned: int, n_local: int):
        _l.check(_h.b200host_blas_reserve_allocation(self.ptr, n_owned, n_local), "blas_reserve_allocation")

    def dot(self, n: int, x, y) -> float:
        return _h.b200host_blas_dot(self.ptr, n, _p(x), _p(y))

    def set_allocator(self, alloc, free):
        """alloc(nbytes) -> device address of zero-filled memory, free(address): used for every vector a solver
        allocates (work vectors with room for the ghost block / in symmetric memory for distributed solves)."""
        a = C.CFUNCTYPE(vp, vp, C.c_size_t)(lambda _c, n: alloc(int(n)))
        f = C.CFUNCTYPE(None, vp, vp)(lambda _c, p: free(p))
        self.keep += [a, f, alloc, free]
        _h.b200host_blas_set_allocator(self.ptr, a, f, None)


class Communicator(Handle):
    """b200fem::Communicator (NCCL, one process per GPU). `broadcast_bytes(bytes | None) -> bytes` is the launcher's way
    of handing rank 0's unique id to every rank (a torch.distributed broadcast under torchrun, MPI_Bcast elsewhere)."""

    def __init__(self, world: int, rank: int, broadcast_bytes=None):
        uid = None
        if world > 1:
            buf = C.create_string_buffer(_l.UNIQUE_ID_BYTES)
            if rank == 0:
                _l.check(_h.b200host_comm_unique_id(buf), "comm_unique_id")
            uid = broadcast_bytes(bytes(buf.raw) if rank == 0 else None)
            assert len(uid) == _l.UNIQUE_ID_BYTES
        self.world, self.rank = world, rank
        super().__init__(_h.b200host_comm(world, rank, uid))

    def sum(self, v: float) -> float:
        return _h.b200host_comm_sum(self.ptr, float(v))

    def max(self, v: float) -> float:
        return _h.b200host_comm_max(self.ptr, float(v))

    def barrier(self):
        _h.b200host_comm_barrier(self.ptr)

    def error(self) -> int:
        return _h.b200host_comm_error(self.ptr)


def torch_broadcast_bytes(group=None):
    """broadcast_bytes for Communicator under torchrun: rank 0's bytes to everyone through torch.distributed."""
    import torch.distributed as dist

    def bc(payload):
        box = [payload]
        dist.broadcast_object_list(box, src=0, group=group)
        return box[0]

    return bc


class Halo(Handle):
    """b200fem::HaloExchange from a plan dict (sfem_b200.parallel.halo_plan); COLLECTIVE over the communicator."""

    def __init__(self, comm: Communicator, dtype, block_size: int, plan: dict, transport: int = 0):
        ci = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        cp = lambda a: np.ascontiguousarray(a, dtype=np.int64)
        sp, so, si = ci(plan["send_peers"]), cp(plan["send_offsets"]), ci(plan["send_idx"])
        rp, rb, rc = ci(plan["recv_peers"]), cp(plan["recv_begin"]), cp(plan["recv_count"])
        self.n_owned, self.n_local, self.bs = int(plan["n_owned"]), int(plan["n_local"]), block_size
        ptr = _h.b200host_halo(comm.ptr, _l._rt(dtype), block_size, self.n_owned, self.n_local, sp.size, sp.ctypes.data,
                               so.ctypes.data, si.ctypes.data if si.size else None, rp.size, rp.ctypes.data, rb.ctypes.data,
                               rc.ctypes.data, transport)
        super().__init__(ptr, keep=[comm])
        if "recv_idx" in plan:  # indexed receive: only the ghosts that can reach an owned row travel
            ro, ri = cp(plan["recv_offsets"]), ci(plan["recv_idx"])
            _l.check(_h.b200host_halo_set_recv_index(self.ptr, ro.ctypes.data, ri.ctypes.data if ri.size else None), "halo_set_recv_index")

    def gather(self, x_local):
        _l.check(_h.b200host_halo_gather(self.ptr, _p(x_local)), "halo_gather")

    @property
    def transport(self) -> str:
        return {1: "push", 2: "nccl"}.get(_h.b200host_halo_transport(self.ptr), "?")

    @property
    def values_sent(self) -> int:
        return _h.b200host_halo_values_sent(self.ptr)

    @property
    def launches_per_gather(self) -> int:
        return _h.b200host_halo_launches_per_gather(self.ptr)


def parallel_op(local_op: Handle, in_halo: Halo, owned_rows: int, owned_cols: int, independent_range=None) -> Handle:
    """ParallelMatrixFreeOperator: gather the ghost block of x (input space), then the local operator. With
    independent_range = (e0, e1) (elements that read no ghost value) the overwriting apply runs them between the two
    halves of the gather; .overlapped tells whether the local operator supports it."""
    h = Handle(_h.b200host_parallel_op(local_op.ptr, in_halo.ptr, owned_rows, owned_cols), keep=[local_op, in_halo])
    h.overlapped = False
    if independent_range is not None and independent_range[1] > independent_range[0]:
        h.overlapped = _h.b200host_parallel_op_set_independent_range(h.ptr, int(independent_range[0]), int(independent_range[1])) == 0
    return h


class Function(Handle):
    """Operators (lib.Op plans) + Dirichlet conditions on one space; is itself the Operator of make_linear_op(f)."""

    def __init__(self, dtype, n_nodes: int, block_size: int = 1):
        import torch

        self.rt, self.n_nodes, self.block_size, self.dtype = _l._rt(dtype), n_nodes, block_size, dtype
        self._torch = torch
        super().__init__(_h.b200host_function(self.rt, n_nodes, block_size))

    @property
    def n_dofs(self):
        return self.n_nodes * self.block_size

    def add_operator(self, op: _l.Op):
        self.keep.append(op)
        _h.b200host_function_add_operator(self.ptr, op.handle)

    def add_dirichlet(self, nodes, component: int, value: float):
        idx = self._torch.as_tensor(np.ascontiguousarray(nodes, dtype=np.int32)).cuda()
        self.keep.append(idx)
        _h.b200host_function_add_dirichlet(self.ptr, _p(idx), idx.numel(), component, value)

    def apply_constraints(self, x):
        _l.check(_h.b200host_function_apply_constraints(self.ptr, _p(x)), "apply_constraints")

    def apply_zero_constraints(self, x):
        _l.check(_h.b200host_function_apply_zero_constraints(self.ptr, _p(x)), "apply_zero_constraints")

    def set_value_to_constrained_dofs(self, v, x):
        _l.check(_h.b200host_function_set_value_to_constrained_dofs(self.ptr, v, _p(x)), "set_value_to_constrained_dofs")

    def hessian_diag(self, d):
        _l.check(_h.b200host_function_hessian_diag(self.ptr, _p(d)), "hessian_diag")

    def hessian_block_diag_sym(self, d):
        _l.check(_h.b200host_function_hessian_block_diag_sym(self.ptr, _p(d)), "hessian_block_diag_sym")

    def launches_per_apply(self, overwrite=True) -> int:
        return _h.b200host_function_launches_per_apply(self.ptr, int(overwrite))


def callback_op(dtype, rows, cols, apply, overwrite=None) -> Handle:
    """Operator from Python callables taking raw device addresses: apply(x_addr, y_addr) accumulates."""
    a = APPLY_CB(lambda _c, x, y: int(apply(x, y) or 0))
    o = APPLY_CB(lambda _c, x, y: int(overwrite(x, y) or 0)) if overwrite else C.cast(None, APPLY_CB)
    return Handle(_h.b200host_callback_op(_l._rt(dtype), rows, cols, a, o, None), keep=[a, o, apply, overwrite])


def restriction(dtype, transfer: _l.Transfer, vec_size=1, coarse_function: Function | None = None) -> Handle:
    return Handle(_h.b200host_restriction(_l._rt(dtype), transfer.handle, vec_size,
                                          coarse_function.ptr if coarse_function else None), keep=[transfer, coarse_function])


def prolongation(dtype, transfer: _l.Transfer, vec_size=1, fine_function: Function | None = None) -> Handle:
    return Handle(_h.b200host_prolongation(_l._rt(dtype), transfer.handle, vec_size,
                                           fine_function.ptr if fine_function else None), keep=[transfer, fine_function])


def cg(blas: Blas, op: Handle, precond: Handle | None = None, rtol=1e-10, atol=1e-16, max_it=10000) -> Handle:
    return Handle(_h.b200host_cg(blas.ptr, op.ptr, precond.ptr if precond else None, rtol, atol, max_it),
                  keep=[blas, op, precond])


def jacobi(blas: Blas, diag, relaxation: float) -> Handle:
    return Handle(_h.b200host_jacobi(blas.ptr, _p(diag), diag.numel(), relaxation), keep=[blas, diag])


def block_sym_jacobi(blas: Blas, diag6, mask, relaxation: float) -> Handle:
    return Handle(_h.b200host_block_sym_jacobi(blas.ptr, _p(diag6), diag6.numel() // 6, _p(mask), relaxation),
                  keep=[blas, diag6, mask])


def mixed_block_sym_jacobi(diag6, mask, relaxation: float) -> Handle:
    """MixedPrecisionShiftableBlockSymJacobi<double, float>: fp64 block diagonal in, fp32 inverted blocks kept."""
    return Handle(_h.b200host_mixed_block_sym_jacobi(_p(diag6), diag6.numel() // 6, _p(mask), relaxation), keep=[diag6, mask])


def stationary(blas: Blas, op: Handle, precond: Handle, max_it: int) -> Handle:
    return Handle(_h.b200host_stationary(blas.ptr, op.ptr, precond.ptr, max_it), keep=[blas, op, precond])


def cheb3(blas: Blas, op: Handle, eig_tol=1e-6, eig_max_it=1000, scale_eig_max=1.0, scale_eig_min=0.06, max_it=3,
          initial_guess_zero=False, init="ones") -> Handle:
    h = Handle(_h.b200host_cheb3(blas.ptr, op.ptr, eig_tol, eig_max_it, scale_eig_max, scale_eig_min, max_it,
                                 int(initial_guess_zero)), keep=[blas, op])
    if isinstance(init, str) and init == "ones":
        _h.b200host_cheb3_init_with_ones(h.ptr)
    elif init is not None:
        _h.b200host_cheb3_init(h.ptr, _p(init))  # device tensor: the power method's start vector
    return h


def cheb3_eig_max(h: Handle) -> float:
    return _h.b200host_cheb3_eig_max(h.ptr)


def cheb3_power_method_iterations(h: Handle) -> int:
    return _h.b200host_cheb3_power_method_iterations(h.ptr)


def cheb3_set_initial_guess_zero(h: Handle, val: bool):
    _h.b200host_cheb3_set_initial_guess_zero(h.ptr, int(val))


def mg(blas: Blas, max_it=10, atol=1e-10, cycle_type=1, verbose=False) -> Handle:
    return Handle(_h.b200host_mg(blas.ptr, max_it, atol, cycle_type, int(verbose)), keep=[blas])


def mg_add_level(m: Handle, op, smoother: Handle, prolong, restrict):
    p = lambda h: h.ptr if h is not None else None
    _h.b200host_mg_add_level(m.ptr, p(op), smoother.ptr, p(prolong), p(restrict))
    m.keep += [op, smoother, prolong, restrict]


def apply(obj: Handle, b, x) -> int:
    """Solvers: solve A x = b in place in x. Operators: x += A b. Returns the reference's int code (0 = success)."""
    return _h.b200host_apply(obj.ptr, _p(b), _p(x))


def apply_overwrite(obj: Handle, x, y) -> int:
    return _h.b200host_apply_overwrite(obj.ptr, _p(x), _p(y))


def iterations(obj: Handle) -> int:
    return _h.b200host_iterations(obj.ptr)


def residual_history(obj: Handle) -> np.ndarray:
    n = _h.b200host_residual_history(obj.ptr, None, 0)
    out = np.zeros(max(n, 0))
    if n > 0:
        _h.b200host_residual_history(obj.ptr, out.ctypes.data_as(C.POINTER(f64)), n)
    return out


def constraints_mask(n_dofs: int, block_size: int, conditions) -> np.ndarray:
    """Bit mask of constrained dofs in the reference's mask_t layout (algebra/sfem_mask.hpp:13-44; built by
    Function::constraints_mask): bit (i % 8) of byte (i / 8) for dof i = node * block_size + component."""
    m = np.zeros((n_dofs + 7) // 8, dtype=np.uint8)
    for nodes, comp, _v in conditions:
        dof = np.asarray(nodes, dtype=np.int64) * block_size + comp
        np.bitwise_or.at(m, dof // 8, (1 << (dof % 8)).astype(np.uint8))
    return m


def read_mesh(folder: str) -> dict:
    """A mesh folder in the reference's on-disk format (x/y/z.raw fp32, i<k>.raw int32, meta.yaml) through the C++ reader
    (sfem_b200/host/b200fem_mesh_io.hpp): {"element_type", "elements" [nxe][ne] int32, "points" [dim][nn] fp32}."""
    m = _h.b200host_mesh_read(folder.encode())
    if not m:
        raise _l.B200FemError("read_mesh: " + _h.b200host_mesh_last_error().decode())
    try:
        ne, nn = _h.b200host_mesh_n_elements(m), _h.b200host_mesh_n_nodes(m)
        nxe, dim = _h.b200host_mesh_nxe(m), _h.b200host_mesh_dim(m)
        if ne:
            el = np.stack([np.ctypeslib.as_array(_h.b200host_mesh_elements(m, k), shape=(ne,)).copy() for k in range(nxe)])
        else:
            el = np.zeros((nxe, 0), np.int32)
        pts = np.stack([np.ctypeslib.as_array(_h.b200host_mesh_points(m, d), shape=(nn,)).copy() for d in range(dim)])
        return {"element_type": _h.b200host_mesh_element_type(m).decode(), "elements": el, "points": pts}
    finally:
        _h.b200host_mesh_destroy(m)


def write_mesh(folder: str, element_type: str, elements: np.ndarray, points: np.ndarray):
    os.makedirs(folder, exist_ok=True)
    el = np.ascontiguousarray(elements, dtype=np.int32)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    rc = _h.b200host_mesh_write(folder.encode(), element_type.encode(), el.shape[0], el.shape[1], pts.shape[1], pts.shape[0],
                                _l._soa_ptrs(el, C.c_int32), _l._soa_ptrs(pts, C.c_float))
    if rc:
        raise _l.B200FemError("write_mesh: " + _h.b200host_mesh_last_error().decode())

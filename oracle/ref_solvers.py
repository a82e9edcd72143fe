"""ctypes front for oracle/_ref/libsfem_ref_solvers_{f64,f32}.so. TEST INFRASTRUCTURE ONLY.

The library is the reference's own header-only solvers (algebra/sfem_cg.hpp, sfem_Chebyshev3.hpp,
sfem_PowerMethod.hpp, sfem_Stationary.hpp, sfem_ShiftableJacobi.hpp, sfem_Multigrid.hpp) instantiated for real_t by
oracle/ref_shim/ref_solvers.cpp. Operators are Python callables wrapped as sfem::LambdaOperator:

* execution space HOST: the callable receives numpy views of the solver's host vectors (`fn(x, y)`, y accumulates);
  BLAS is the reference's OpenMP_BLAS. This is the oracle for iteration counts.
* execution space DEVICE: the callable receives raw device addresses (ints); BLAS is bound to the b200fem C ABI with
  `set_device_table`. This is the drop-in check: the reference's solver code driving our kernels.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST, DEVICE = 0, 1

DEVICE_TABLE_ORDER = [
    "b200fem_malloc", "b200fem_free", "b200fem_blas_zeros", "b200fem_blas_values", "b200fem_blas_copy",
    "b200fem_blas_scal", "b200fem_blas_axpy", "b200fem_blas_axpby", "b200fem_blas_zaxpby", "b200fem_blas_xypaz",
    "b200fem_blas_reciprocal", "b200fem_blas_dot", "b200fem_blas_norm2",
]


def lib_path(real: str) -> str:
    return os.path.join(HERE, "_ref", f"libsfem_ref_solvers_{real}.so")


def available(real: str = "f64") -> bool:
    return os.path.exists(lib_path(real))


class Handle:
    def __init__(self, owner, ptr, keep=()):
        self.owner, self.ptr, self.keep = owner, ptr, list(keep)

    def __del__(self):
        try:
            if self.ptr:
                self.owner.lib.refsolv_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class RefSolvers:
    def __init__(self, real: str = "f64"):
        path = lib_path(real)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(path)
        self.real = real
        self.rt = np.float64 if real == "f64" else np.float32
        self.crt = C.c_double if real == "f64" else C.c_float
        L = self.lib
        self.APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
        L.refsolv_make_op.restype = C.c_void_p
        L.refsolv_make_op.argtypes = [C.c_ssize_t, C.c_ssize_t, self.APPLY_FN, C.c_void_p, C.c_int]
        L.refsolv_cg.restype = C.c_void_p
        L.refsolv_cg.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
        L.refsolv_cg_set_host_fusion.argtypes = [C.c_void_p, C.c_int]
        L.refsolv_jacobi.restype = C.c_void_p
        L.refsolv_jacobi.argtypes = [C.c_void_p, C.c_ssize_t, C.c_double, C.c_int]
        L.refsolv_block_sym_jacobi.restype = C.c_void_p
        L.refsolv_block_sym_jacobi.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_double]
        L.refsolv_stationary.restype = C.c_void_p
        L.refsolv_stationary.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.refsolv_cheb3.restype = C.c_void_p
        L.refsolv_cheb3.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
        L.refsolv_cheb3_init_with_ones.argtypes = [C.c_void_p]
        L.refsolv_cheb3_init.argtypes = [C.c_void_p, C.c_void_p]
        L.refsolv_cheb3_eig_max.restype = C.c_double
        L.refsolv_cheb3_eig_max.argtypes = [C.c_void_p]
        L.refsolv_cheb3_set_initial_guess_zero.argtypes = [C.c_void_p, C.c_int]
        L.refsolv_mg.restype = C.c_void_p
        L.refsolv_mg.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        L.refsolv_mg_add_level.argtypes = [C.c_void_p] * 5
        L.refsolv_apply.restype = C.c_int
        L.refsolv_apply.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.refsolv_iterations.restype = C.c_int
        L.refsolv_iterations.argtypes = [C.c_void_p]
        L.refsolv_destroy.argtypes = [C.c_void_p]
        L.refsolv_power_method.restype = C.c_double
        L.refsolv_power_method.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_int]
        L.refsolv_set_device_table.argtypes = [C.c_void_p]
        assert L.refsolv_sizeof_real() == np.dtype(self.rt).itemsize

    # ---- device binding
    def set_device_table(self, b200fem_cdll):
        """Bind sfem::BLAS<real_t> to the vector entry points of libb200fem.so (function addresses only)."""
        arr = (C.c_void_p * len(DEVICE_TABLE_ORDER))()
        for i, name in enumerate(DEVICE_TABLE_ORDER):
            arr[i] = C.cast(getattr(b200fem_cdll, name), C.c_void_p).value
        self._table = arr
        self.lib.refsolv_set_device_table(arr)

    # ---- operators
    def make_op(self, fn, rows: int, cols: int | None = None, es: int = HOST) -> Handle:
        """HOST: fn(x: ndarray[cols], y: ndarray[rows]) accumulates into y. DEVICE: fn(x_addr: int, y_addr: int)."""
        cols = rows if cols is None else cols
        rt = self.rt

        def trampoline(_ctx, xp, yp):
            if es == HOST:
                x = np.ctypeslib.as_array(C.cast(xp, C.POINTER(self.crt)), shape=(cols,))
                y = np.ctypeslib.as_array(C.cast(yp, C.POINTER(self.crt)), shape=(rows,))
                fn(x, y)
            else:
                fn(xp, yp)
            return 0

        cb = self.APPLY_FN(trampoline)
        return Handle(self, self.lib.refsolv_make_op(rows, cols, cb, None, es), keep=[cb, fn, rt])

    def cg(self, op: Handle, precond: Handle | None = None, rtol=1e-10, atol=1e-16, max_it=10000, es=HOST) -> Handle:
        return Handle(self, self.lib.refsolv_cg(op.ptr, precond.ptr if precond else None, rtol, atol, max_it, es),
                      keep=[op, precond])

    def cg_set_verbose(self, cg: Handle, val: bool):
        """the reference's own monitor() output (one line per iteration on stdout); pair with capture_stdout()"""
        self.lib.refsolv_cg_set_verbose.argtypes = [C.c_void_p, C.c_int]
        self.lib.refsolv_cg_set_verbose.restype = None
        self.lib.refsolv_cg_set_verbose(cg.ptr, int(val))

    def cg_set_host_fusion(self, cg: Handle, val: bool):
        self.lib.refsolv_cg_set_host_fusion(cg.ptr, int(val))

    def jacobi(self, diag, relaxation: float, es=HOST, n=None) -> Handle:
        """diag: numpy array (HOST) or device address (DEVICE, pass n)."""
        if es == HOST:
            diag = np.ascontiguousarray(diag, dtype=self.rt)
            return Handle(self, self.lib.refsolv_jacobi(diag.ctypes.data, diag.size, relaxation, es), keep=[diag])
        return Handle(self, self.lib.refsolv_jacobi(diag, n, relaxation, es))

    def block_sym_jacobi(self, diag6, mask: np.ndarray, relaxation: float) -> Handle:
        diag6 = np.ascontiguousarray(diag6, dtype=self.rt)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        return Handle(self, self.lib.refsolv_block_sym_jacobi(diag6.ctypes.data, diag6.size // 6, mask.ctypes.data,
                                                              mask.size, relaxation), keep=[diag6, mask])

    def mixed_block_sym_jacobi(self, diag6, mask: np.ndarray, relaxation: float) -> Handle:
        """MixedPrecisionShiftableBlockSymJacobi<double, float> (f64 build only)"""
        diag6 = np.ascontiguousarray(diag6, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        fn = self.lib.refsolv_mixed_block_sym_jacobi
        fn.restype = C.c_void_p
        fn.argtypes = [C.c_void_p, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_double]
        return Handle(self, fn(diag6.ctypes.data, diag6.size // 6, mask.ctypes.data, mask.size, relaxation), keep=[diag6, mask])

    def stationary(self, op: Handle, precond: Handle, max_it: int, es=HOST) -> Handle:
        return Handle(self, self.lib.refsolv_stationary(op.ptr, precond.ptr, max_it, es), keep=[op, precond])

    def cheb3(self, op: Handle, eig_tol=1e-6, eig_max_it=1000, scale_eig_max=1.0, scale_eig_min=0.06, max_it=3,
              initial_guess_zero=False, es=HOST, init="ones") -> Handle:
        h = Handle(self, self.lib.refsolv_cheb3(op.ptr, eig_tol, eig_max_it, scale_eig_max, scale_eig_min, max_it,
                                                int(initial_guess_zero), es), keep=[op])
        if isinstance(init, str) and init == "ones":
            self.lib.refsolv_cheb3_init_with_ones(h.ptr)
        elif init is not None:  # numpy start vector (HOST) or device address (DEVICE)
            self.lib.refsolv_cheb3_init(h.ptr, init.ctypes.data if isinstance(init, np.ndarray) else init)
        return h

    def cheb3_eig_max(self, h: Handle) -> float:
        return self.lib.refsolv_cheb3_eig_max(h.ptr)

    def cheb3_set_initial_guess_zero(self, h: Handle, val: bool):
        self.lib.refsolv_cheb3_set_initial_guess_zero(h.ptr, int(val))

    def mg(self, max_it=10, atol=1e-10, cycle_type=1, es=HOST, verbose=False) -> Handle:
        return Handle(self, self.lib.refsolv_mg(es, max_it, atol, cycle_type, int(verbose)))

    def mg_add_level(self, mg: Handle, op, smoother: Handle, prolongation, restriction):
        p = lambda h: h.ptr if h is not None else None
        self.lib.refsolv_mg_add_level(mg.ptr, p(op), smoother.ptr, p(prolongation), p(restriction))
        mg.keep += [op, smoother, prolongation, restriction]

    # ---- run
    def apply(self, obj: Handle, b, x) -> int:
        """HOST: numpy arrays of real_t (x updated in place). DEVICE: device addresses."""
        if isinstance(b, np.ndarray):
            assert b.dtype == self.rt and x.dtype == self.rt and b.flags.c_contiguous and x.flags.c_contiguous
            return self.lib.refsolv_apply(obj.ptr, b.ctypes.data, x.ctypes.data)
        return self.lib.refsolv_apply(obj.ptr, b, x)

    def iterations(self, obj: Handle) -> int:
        return self.lib.refsolv_iterations(obj.ptr)

    def power_method(self, op: Handle, max_it: int, tol: float, eigen_vector: np.ndarray, work: np.ndarray) -> float:
        return self.lib.refsolv_power_method(op.ptr, max_it, tol, eigen_vector.size, eigen_vector.ctypes.data,
                                             work.ctypes.data, HOST)


class capture_stdout:
    """fd-level capture of what the reference's solvers printf (their residual monitors): `with capture_stdout() as c: ...;
    c.text` afterwards."""

    def __enter__(self):
        import os
        import sys
        import tempfile

        sys.stdout.flush()
        C.CDLL(None).fflush(None)
        self._saved = os.dup(1)
        self._tmp = tempfile.TemporaryFile(mode="w+b")
        os.dup2(self._tmp.fileno(), 1)
        self.text = ""
        return self

    def __exit__(self, *exc):
        import os

        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        self._tmp.seek(0)
        self.text = self._tmp.read().decode(errors="replace")
        self._tmp.close()
        return False


def parse_cg_monitor(text: str):
    """[(iteration, abs residual)] from the reference's CG monitor lines."""
    import re

    return [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"^(\d+): residual abs: ([-+0-9.eEinfa]+)", text, flags=re.M)]


def parse_mg_monitor(text: str):
    """[abs residual per cycle] from the reference's Multigrid monitor lines (sfem_Multigrid.hpp:274-296)."""
    import re

    return [float(m.group(2)) for m in re.finditer(r"^(\d+)\t([-+0-9.eE]+)", text, flags=re.M)]

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/b200fem.h declares, and fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes as C

import numpy as np
import pytest


def test_library_exports_every_declared_symbol():
    from sfem_b200 import lib

    names = lib.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib.raw(), n)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sfem_b200 import lib, mesh

    el, pts = mesh.box_tet4(3, 3, 3)
    fff = np.ones(el.shape[1] * 6, dtype=np.float32)
    with pytest.raises(lib.B200FemError, match="no CUDA device|CUDA error"):
        lib.laplacian(lib.TET4, el, fff, pts.shape[1], np.float64)
    with pytest.raises(lib.B200FemError):
        lib.fff_fill(lib.TET4, el, pts)


def test_bad_arguments_are_reported():
    from sfem_b200 import lib

    h = C.c_void_p()
    rc = lib.raw().b200fem_laplacian_create(C.byref(h), lib.TET4, 0, 1, 1, None, None, 3)
    assert rc == 1 and lib.raw().b200fem_last_error()
    assert lib.raw().b200fem_op_apply(None, None, None, None) == 1


def test_host_layer_library_exports():
    """libb200fem_host.so (C++ host layer behind C entry points) loads and exports what sfem_b200/host.py binds."""
    from sfem_b200 import host

    missing = [n for n in host.exported_symbols() if not hasattr(host.raw(), n)]
    assert not missing, missing
    # object construction needs no device; solving does (no CPU fallback: allocation fails without a GPU)
    import torch

    b = host.Blas(torch.float64)
    assert b.launches() == 0

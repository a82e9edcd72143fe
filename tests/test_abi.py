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
    # the section-8f entry points: no CPU path behind any of them either (the packer is host code and does run)
    el10, pts10 = mesh.tet4_to_tet10(el, pts)
    n10 = pts10.shape[1]
    adj, det = np.ones(el.shape[1] * 9, dtype=np.float32), np.ones(el.shape[1], dtype=np.float32)
    pm = lib.PackedMesh(el10, n10, 16)
    assert pm.n_packs == (el10.shape[1] + 15) // 16
    for make in (lambda: lib.laplacian(lib.TET10, el10, fff, n10, np.float64),
                 lambda: lib.linear_elasticity(lib.MACRO_TET4, el10, adj, det, 1.0, 1.0, n10, np.float64),
                 lambda: lib.laplacian_from_points(lib.TET10, el10, pts10, n10, np.float64),
                 lambda: lib.linear_elasticity_from_points(lib.MACRO_TET4, el10, pts10, 1.0, 1.0, n10, np.float64),
                 lambda: lib.mass(lib.TET10, el10, pts10, n10, np.float64),
                 lambda: lib.packed_laplacian(lib.TET10, pm, fff, np.float64),
                 lambda: lib.sshex8_lumped_mass(1, mesh.box_hex8(3, 3, 3)[0], mesh.box_hex8(3, 3, 3)[1], 27, np.float64)):
        with pytest.raises(lib.B200FemError, match="no CUDA device|CUDA error"):
            make()


def test_new_entry_points_reject_bad_arguments():
    from sfem_b200 import lib

    h = C.c_void_p()
    raw = lib.raw()
    assert raw.b200fem_packed_laplacian_create(C.byref(h), lib.TET4, 1, 4, 4, 8, 8, None, None, None, None, None, None, 8) == 1
    assert b"NULL" in raw.b200fem_last_error()
    assert raw.b200fem_packed_laplacian_set_two_pass(None, 0, None, None, None) == 1
    assert raw.b200fem_packed_mesh_create(C.byref(h), 4, 3, 5, None, 0) == 1
    assert raw.b200fem_sshex8_vector_laplacian_create(C.byref(h), 3, 0, 0, None, None, 3, 8) == 1  # level 3 is not built
    assert raw.b200fem_sshex8_lumped_mass_create(C.byref(h), 2, 0, 0, None, None, 8) == 1  # points is NULL


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

"""CPU: the reference's on-disk mesh format through the C++ reader / writer of the host layer
(sfem_b200/host/b200fem_mesh_io.hpp). tests/golden/ref_mesh/box_{tet4,hex8} were written by the REFERENCE's own generator
(python/sfem/mesh/box_mesh.py -c tet4|hex8 -x 4 -y 3 -z 5 --skip_boundary=1, run in the build container): reading them
pins the format, and comparing them with sfem_b200.mesh.box_tet4 / box_hex8 pins our restatement of that generator
(node numbering x*nz + y*nz*nx + z, hex -> 6 tets, VTK corner order) which every other fixture of the test-suite builds on."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_mesh")


def _host():
    pytest.importorskip("torch")
    try:
        from sfem_b200 import host
    except ImportError:
        pytest.skip("host layer not built")
    return host


@pytest.mark.parametrize("kind", ["tet4", "hex8"])
def test_reads_the_reference_generators_folders(kind):
    host = _host()
    from sfem_b200 import mesh

    m = host.read_mesh(os.path.join(GOLD, f"box_{kind}"))
    assert m["element_type"] == kind.upper()
    el, pts = (mesh.box_tet4 if kind == "tet4" else mesh.box_hex8)(4, 3, 5)
    assert m["elements"].dtype == np.int32 and m["points"].dtype == np.float32
    assert m["points"].shape == pts.shape and np.array_equal(m["points"], pts), "box coordinates differ from the reference generator's"
    assert m["elements"].shape == el.shape
    if kind == "hex8":
        assert np.array_equal(m["elements"], el), "hex8 connectivity differs from the reference generator's"
    else:
        # the same tets: equal as a set of node quadruples in the reference's corner order
        ours = {tuple(c) for c in el.T.tolist()}
        theirs = {tuple(c) for c in m["elements"].T.tolist()}
        assert ours == theirs, "tet4 connectivity differs from the reference generator's"


def test_round_trip_and_errors(tmp_path):
    host = _host()
    from sfem_b200 import lib, mesh

    el, pts = mesh.box_hex8(3, 4, 2)
    pts = mesh.warp_points(pts)
    folder = str(tmp_path / "m")
    host.write_mesh(folder, "HEX8", el, pts)
    assert sorted(os.listdir(folder)) == sorted([f"i{k}.raw" for k in range(8)] + ["x.raw", "y.raw", "z.raw", "meta.yaml"])
    m = host.read_mesh(folder)
    assert m["element_type"] == "HEX8" and np.array_equal(m["elements"], el) and np.array_equal(m["points"], pts)
    # the pattern-based read of the reference's C++ side (no meta.yaml)
    os.remove(os.path.join(folder, "meta.yaml"))
    m2 = host.read_mesh(folder)
    assert np.array_equal(m2["elements"], el) and np.array_equal(m2["points"], pts) and m2["element_type"] == "HEX8"
    # a connectivity entry outside the coordinate arrays is refused
    bad = el.copy()
    bad[3, 1] = pts.shape[1] + 5
    host.write_mesh(folder, "HEX8", bad, pts)
    with pytest.raises(lib.B200FemError):
        host.read_mesh(folder)
    with pytest.raises(lib.B200FemError):
        host.read_mesh(str(tmp_path / "does_not_exist"))


def test_tet10_folder_round_trip(tmp_path):
    """10-node tetrahedra travel through the same folder format (i0.raw .. i9.raw); without meta.yaml 10 index files read as TET10
    (MACRO_TET4 is the same connectivity under another element_type string)."""
    host = _host()
    from sfem_b200 import mesh

    el4, pts4 = mesh.box_tet4(3, 4, 3)
    el, pts = mesh.tet4_to_tet10(el4, mesh.warp_points(pts4))
    assert el.shape[0] == 10 and np.array_equal(el[:4], el4)
    # mid-edge nodes sit in the middle of their edge, every edge has one node
    p = pts.astype(np.float64)
    for k, (a, b) in enumerate(mesh.TET10_EDGES):
        assert np.abs(p[:, el[4 + k]] - 0.5 * (p[:, el[a]] + p[:, el[b]])).max() < 1e-6
    edges = {tuple(sorted((int(el[a, e]), int(el[b, e])))) for a, b in mesh.TET10_EDGES for e in range(el.shape[1])}
    assert pts.shape[1] == pts4.shape[1] + len(edges)
    for et in ("TET10", "MACRO_TET4"):
        folder = str(tmp_path / et)
        host.write_mesh(folder, et, el, pts)
        m = host.read_mesh(folder)
        assert m["element_type"] == et and np.array_equal(m["elements"], el) and np.array_equal(m["points"], pts)
    os.remove(os.path.join(str(tmp_path / "TET10"), "meta.yaml"))
    assert host.read_mesh(str(tmp_path / "TET10"))["element_type"] == "TET10"

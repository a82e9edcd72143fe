"""b200fem_packed_mesh_create (host code inside libb200fem.so, no GPU needed): the arrays it returns satisfy the contract
PackedLaplacian::apply relies on (frontend/ops/sfem_PackedLaplacian.cpp:111-176): owned ranges partition the nodes, the last
n_shared of a pack's range are exactly the nodes other packs touch too (everything before them takes a plain `+=` without
atomics), ghosts belong to other packs, 16-bit local indices stay inside the pack, and expanding the packed connectivity
gives back the input mesh under the node permutation."""
import numpy as np
import pytest


def _meshes():
    from sfem_b200 import mesh

    el4, pts4 = mesh.box_tet4(7, 6, 8)
    el8, _ = mesh.box_hex8(6, 7, 5)
    el10, pts10 = mesh.tet4_to_tet10(el4, pts4)
    return {"tet4": (el4, pts4.shape[1]), "hex8": (el8, 6 * 7 * 5), "tet10": (el10, pts10.shape[1])}


@pytest.mark.parametrize("kind", ["tet4", "hex8", "tet10"])
@pytest.mark.parametrize("epp", [0, 64, 100, 5000])
def test_packed_mesh_contract(kind, epp):
    from sfem_b200 import lib

    el, nn = _meshes()[kind]
    ne = el.shape[1]
    pm = lib.PackedMesh(el, nn, epp)
    E = pm.n_elements_per_pack
    assert E == (1024 if epp == 0 else epp) and pm.n_packs == (ne + E - 1) // E
    # the permutation is a permutation, the owned ranges partition [0, nn)
    assert np.array_equal(np.sort(pm.new_to_old), np.arange(nn))
    assert pm.owned_nodes_ptr[0] == 0 and pm.owned_nodes_ptr[-1] == nn and np.all(np.diff(pm.owned_nodes_ptr) >= 0)
    # expanding the local indices gives the input connectivity back
    g = pm.global_elements()
    assert np.array_equal(pm.new_to_old[g], el)
    # which packs touch a node
    pack = np.arange(ne) // E
    touched = {}
    for v in range(g.shape[0]):
        for n, p in zip(g[v], pack):
            touched.setdefault(int(n), set()).add(int(p))
    max_local = 0
    for p in range(pm.n_packs):
        b, e = pm.owned_nodes_ptr[p], pm.owned_nodes_ptr[p + 1]
        n_excl = (e - b) - pm.n_shared[p]
        for n in range(b, e):
            packs = touched.get(n, set())
            if n < b + n_excl:
                assert packs <= {p}, "an exclusively owned node is touched by another pack"
            else:
                assert p in packs and len(packs) > 1, "a node in the shared tail is not shared"
        gh = pm.ghost_idx[pm.ghost_ptr[p]:pm.ghost_ptr[p + 1]]
        assert np.all((gh < b) | (gh >= e)) and np.all(np.diff(gh) > 0)
        for n in gh:
            assert p in touched[int(n)]
        n_local = (e - b) + gh.size
        max_local = max(max_local, n_local)
        sl = slice(p * E, min(ne, (p + 1) * E))
        assert pm.elements[:, sl].max() < n_local
    assert pm.max_nodes_per_pack == max_local <= 65536
    # the two-pass lists: every ghost entry is reduced exactly once, into the node it stands for, one row per shared ghost node
    assert pm.ghost_reduce_ptr[0] == 0 and pm.ghost_reduce_ptr[-1] == pm.n_ghost_entries
    assert np.array_equal(np.sort(pm.ghost_reduce_idx), np.arange(pm.n_ghost_entries))
    assert np.all(np.diff(pm.ghost_reduce_dest) > 0) and np.array_equal(pm.ghost_reduce_dest, np.unique(pm.ghost_idx))
    rows = np.repeat(np.arange(pm.ghost_reduce_dest.size), np.diff(pm.ghost_reduce_ptr))
    assert np.array_equal(pm.ghost_idx[pm.ghost_reduce_idx], pm.ghost_reduce_dest[rows])


def test_packed_mesh_rejects_bad_input():
    from sfem_b200 import lib

    el = np.array([[0], [1], [2], [7]], dtype=np.int32)
    with pytest.raises(RuntimeError):
        lib.PackedMesh(el, 4, 0)


def test_packed_mesh_edge_cases():
    """Empty mesh, a ragged last pack, nodes no element touches (they stay in the owned range of the last pack)."""
    from sfem_b200 import lib

    pm = lib.PackedMesh(np.zeros((4, 0), dtype=np.int32), 5, 0)
    assert pm.n_packs == 0 and pm.n_ghost_entries == 0 and pm.elements.shape == (4, 0) and pm.ghost_reduce_ptr.tolist() == [0]
    # 3 tets, 2 per pack, node 6 untouched
    el = np.array([[0, 1, 2], [1, 2, 3], [2, 3, 4], [3, 4, 5]], dtype=np.int32)
    pm = lib.PackedMesh(el, 7, 2)
    assert pm.n_packs == 2 and pm.owned_nodes_ptr.tolist()[0] == 0 and pm.owned_nodes_ptr[-1] == 7
    assert np.array_equal(pm.new_to_old[pm.global_elements()], el)
    assert pm.new_to_old[-1] == 6  # the orphan node closes the last range
    assert pm.n_shared.tolist() == [3, 0]  # nodes 2, 3, 4 of pack 0 are touched by pack 1 too; pack 1 shares nothing it owns
    assert pm.ghost_ptr.tolist() == [0, 0, 3] and np.array_equal(np.sort(pm.new_to_old[pm.ghost_idx]), [2, 3, 4])

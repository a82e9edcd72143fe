"""Host logic of the distributed apply on CPU: world_size-2/3 gloo processes, the oracle as each rank's local
operator. Mirrors frontend/tests/mpi/sfem_ParallelLaplacianApplyTest.cpp:62-141 of the reference: the
distributed apply must equal the serial apply on the owned rows (1e-10), aura elements must exist, ghost
values must arrive from their owners."""
import os
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case):
    import sys

    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        from oracle.ref import Oracle, available
        from sfem_b200 import mesh, parallel

        orc = Oracle("reference" if available("reference") else "port", "f64")
        if case in ("hex8", "tet4"):
            if case == "hex8":
                el, pts = mesh.box_hex8(7, 6, 9)
            else:
                el, pts = mesh.box_tet4(6, 5, 8)
            pts = mesh.warp_points(pts)
            nn = pts.shape[1]
            u = mesh.quadratic_field(pts)
            if case == "hex8":
                fff = orc.hex8_fff_fill(el, pts)
                ref = orc.hex8_laplacian_apply_opt(el, fff, u)
            else:
                fff = orc.tet4_fff_fill(el, pts)
                ref = orc.tet4_laplacian_apply_opt(el, fff, u)
            part = parallel.partition_elements_by_coordinate(el, pts, world, axis=2)
            lm = parallel.distribute(el, nn, part, rank, world)
            if rank < world - 1:  # interface nodes belong to the lower part, so every part but the last has an aura
                assert lm.elements.shape[1] > lm.n_owned_elements, "no aura elements"
            if world > 1:
                assert lm.n_local > lm.n_owned, "no ghost nodes"
            lpts = pts[:, lm.global_ids]
            lfff = orc.hex8_fff_fill(lm.elements, lpts) if case == "hex8" else orc.tet4_fff_fill(lm.elements, lpts)
            apply = (lambda x: orc.hex8_laplacian_apply_opt(lm.elements, lfff, x)) if case == "hex8" else (
                lambda x: orc.tet4_laplacian_apply_opt(lm.elements, lfff, x))
            u_of_gid = lambda g: u[g]
            ref_of_gid = lambda g: ref[g]
        elif case == "tet10":  # 10-node tetrahedra through the same element partition (the aura rule does not depend on the family)
            if not available("reference"):
                return  # the tet10 operator exists in the reference build of the oracle only
            el4, pts4 = mesh.box_tet4(5, 4, 7)
            el, pts = mesh.tet4_to_tet10(el4, mesh.warp_points(pts4))
            nn = pts.shape[1]
            u = mesh.quadratic_field(pts)
            ref = orc.p2tet_laplacian_apply_opt("tet10", el, orc.tet4_fff_fill(el[:4], pts), u)
            part = parallel.partition_elements_by_coordinate(el[:4], pts, world, axis=2)
            lm = parallel.distribute(el, nn, part, rank, world)
            if rank < world - 1:
                assert lm.elements.shape[1] > lm.n_owned_elements, "no aura elements"
            assert lm.n_local > lm.n_owned, "no ghost nodes"
            lpts = pts[:, lm.global_ids]
            lfff = orc.tet4_fff_fill(lm.elements[:4], lpts)
            apply = lambda x: orc.p2tet_laplacian_apply_opt("tet10", lm.elements, lfff, x)
            u_of_gid = lambda g: u[g]
            ref_of_gid = lambda g: ref[g]
        elif case == "tet4_slab":  # per-rank generation of the strong-scaling tet4 box (config 4) against the one-piece box
            nx, ny, nz = 5, 9, 4
            el, pts = mesh.box_tet4(nx, ny, nz)
            u = mesh.quadratic_field(pts)
            ref = orc.tet4_laplacian_apply_opt(el, orc.tet4_fff_fill(el, pts), u)
            lm = parallel.tet4_slab(nx, ny, nz, rank, world)
            lpts = lm.extra["points"]
            assert np.array_equal(lpts, pts[:, lm.global_ids]), "slab coordinates differ from the one-piece box"
            assert lm.n_owned_elements == 6 * (nx - 1) * (nz - 1) * (np.linspace(0, ny - 1, world + 1).astype(int)[rank + 1]
                                                                      - np.linspace(0, ny - 1, world + 1).astype(int)[rank])
            if rank < world - 1:
                assert lm.elements.shape[1] == lm.n_owned_elements + 6 * (nx - 1) * (nz - 1), "aura layer missing"
            # owned elements reference no node of the layer above the interface plane
            lfff = orc.tet4_fff_fill(lm.elements, lpts)
            apply = lambda x: orc.tet4_laplacian_apply_opt(lm.elements, lfff, x)
            u_of_gid = lambda g: u[g]
            ref_of_gid = lambda g: ref[g]
        else:  # sshex8 slab (weak-scaling fixture) against the serial mesh of the whole stack
            L, mx, my, mz = 4, 3, 2, 2
            lm = parallel.sshex8_slab(L, mx, my, mz, rank, world)
            ser = mesh.sshex8_box(L, mx, my, mz * world, t=float(world))
            gid_ser, _ = parallel.sshex8_box_lattice_keys(L, mx, my, ser["elements"], ser["n_nodes"])
            assert np.unique(gid_ser).size == ser["n_nodes"]
            fff_ser = orc.hex8_fff_fill(ser["m_elements"], ser["m_points"])
            u_ser = np.random.default_rng(7).uniform(-1, 1, ser["n_nodes"])
            ref_ser = orc.sshex8_laplacian_apply(L, ser["elements"], fff_ser, u_ser)
            order = np.argsort(gid_ser)
            look = lambda arr, g: arr[order[np.searchsorted(gid_ser[order], g)]]
            u_of_gid = lambda g: look(u_ser, g)
            ref_of_gid = lambda g: look(ref_ser, g)
            lfff = orc.hex8_fff_fill(lm.extra["m_elements"], lm.extra["m_points"])
            apply = lambda x: orc.sshex8_laplacian_apply(L, lm.elements, lfff, x)
            if rank < world - 1:
                assert lm.elements.shape[1] == mx * my * (mz + 1)

        # every owned node is owned exactly once over the ranks
        cnt = torch.tensor([lm.n_owned], dtype=torch.int64)
        dist.all_reduce(cnt)
        ex = parallel.HaloExchange(lm)
        x_local = torch.zeros(lm.n_local, dtype=torch.float64)
        x_local[: lm.n_owned] = torch.from_numpy(u_of_gid(lm.global_ids[: lm.n_owned]))
        ex.gather(x_local)
        assert np.array_equal(x_local.numpy(), u_of_gid(lm.global_ids)), "ghost values differ from the owners'"
        y_local = apply(x_local.numpy())
        err = np.abs(y_local[: lm.n_owned] - ref_of_gid(lm.global_ids[: lm.n_owned])).max()
        assert err < 1e-10, err
        if case == "sshex8" and world > 1:
            # exchanging only the ghosts that share a micro element with an owned node (b200fem_halo_set_recv_index): the
            # other ghost values cannot reach an owned row -- poison them and compare the owned rows again
            needed = parallel.sshex8_reachable_ghosts(lm)
            assert needed.sum() > 0
            if rank < world - 1:  # ranks with an aura layer: only its first lattice plane can reach an owned row
                assert needed.sum() < needed.size, "the reduced ghost set must be a proper subset of the aura nodes"
            ex2 = parallel.HaloExchange(lm, needed=needed)
            x2 = torch.full((lm.n_local,), 1e30, dtype=torch.float64)
            x2[: lm.n_owned] = x_local[: lm.n_owned]
            ex2.gather(x2)
            got = x2.numpy()[lm.n_owned:]
            assert np.array_equal(got[needed], u_of_gid(lm.global_ids[lm.n_owned:])[needed])
            assert np.all(got[~needed] == 1e30)
            y2 = apply(x2.numpy())
            assert np.abs(y2[: lm.n_owned] - ref_of_gid(lm.global_ids[: lm.n_owned])).max() < 1e-10
        # owned dot product summed over ranks == serial dot product
        d = torch.tensor([float(x_local[: lm.n_owned].numpy() @ y_local[: lm.n_owned])], dtype=torch.float64)
        dist.all_reduce(d)
        tot = int(cnt.item())
        gids = np.arange(tot) if case in ("hex8", "tet4", "tet4_slab", "tet10") else None
        if case == "sshex8":
            # FunctionSpace::derefine of the distributed mesh: the macro corners as their own [owned | ghosts] space,
            # its exchange plan, and a gather on it
            corner_rows = [mesh.sshex8_lidx(L, cx * L, cy * L, cz * L) for cx, cy, cz in mesh.HEX8_CORNERS]
            cm, fine_of_coarse = parallel.coarse_space(lm, corner_rows)
            assert cm.n_owned == int((fine_of_coarse < lm.n_owned).sum()) and cm.n_local == fine_of_coarse.size
            assert np.all(np.diff(np.where(cm.owner == rank, -1, cm.owner)) >= 0), "coarse ghosts not sorted by owner"
            cex = parallel.HaloExchange(cm)
            xc = torch.zeros(cm.n_local, dtype=torch.float64)
            xc[: cm.n_owned] = torch.from_numpy(u_of_gid(cm.global_ids[: cm.n_owned]))
            cex.gather(xc)
            assert np.array_equal(xc.numpy(), u_of_gid(cm.global_ids)), "coarse ghost values differ from the owners'"
            nc = torch.tensor([cm.n_owned], dtype=torch.int64)
            dist.all_reduce(nc)
            assert int(nc.item()) == (mx + 1) * (my + 1) * (mz * world + 1)
        if gids is not None:
            assert abs(d.item() - float(u_of_gid(gids) @ ref_of_gid(gids))) < 1e-9 * max(1.0, abs(d.item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,world", [("hex8", 2), ("tet4", 3), ("sshex8", 2), ("sshex8", 3), ("tet4_slab", 2), ("tet4_slab", 3),
                                        ("tet10", 2), ("tet10", 3)])
def test_distributed_apply_matches_serial(case, world, oracles):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), case), nprocs=world, join=True)

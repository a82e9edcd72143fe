"""Distributed apply: mesh partition, halo plan and the NCCL exchange (one process per GPU).

Mirrors the reference's distributed matrix-free apply (frontend/sfem_ParallelMatrixFreeOperator.cpp:78-105,
tested by frontend/tests/mpi/sfem_ParallelLaplacianApplyTest.cpp:62-141):

* every rank keeps its nodes in the order [owned | ghosts + aura] and every element that touches one of
  its owned nodes (its own elements plus a one-element-deep *aura* of its neighbours' elements);
* apply = copy the owned part of x, pull the ghost/aura values from their owners (`Exchange::gather`),
  run the local operator over all local elements, keep the owned rows. There is no reverse scatter:
  interface rows are completed by redundant computation on the aura layer;
* dot products / norms run over the owned entries only and are summed over the ranks
  (algebra/parallel/sfem_parallel_cg.hpp:169-172).

`external/smesh` (Exchange, distributed Mesh, Communicator) is absent from the reference checkout, so the
partition and the exchange plan are built here from that observable contract. The transport is
`torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests): per apply one grouped
send/recv per neighbour of packed owned values (packed by `b200fem_halo_pack`), received straight into
the contiguous ghost block of the local vector -- ghosts are ordered by owner, so no unpack kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

idx_t = np.int32


@dataclass
class LocalMesh:
    """One rank's share of a mesh, nodes ordered [owned | ghosts sorted by (owner, global id)]."""

    elements: np.ndarray          # [nxe][n_local_elements] int32, local node ids; owned elements first
    n_owned_elements: int
    n_owned: int                  # owned nodes
    n_local: int                  # owned + ghost/aura nodes
    global_ids: np.ndarray        # [n_local] int64
    owner: np.ndarray             # [n_local] int32 rank owning each local node
    rank: int = 0
    world: int = 1
    extra: dict = field(default_factory=dict)


def renumber_owned_first(elements: np.ndarray, global_ids: np.ndarray, owner: np.ndarray, rank: int):
    """Stable renumbering to [owned | ghosts by owner]; keeps the relative order inside every class, so
    entity-wise affine numberings (sshex8 boxes) stay affine."""
    key = np.where(owner == rank, -1, owner).astype(np.int64)
    order = np.argsort(key, kind="stable")          # new position -> old id
    new_of_old = np.empty_like(order)
    new_of_old[order] = np.arange(order.size)
    return new_of_old[elements].astype(idx_t), global_ids[order], owner[order].astype(np.int32), int((owner == rank).sum())


def partition_elements_by_coordinate(elements: np.ndarray, points: np.ndarray, world: int, axis: int = 2) -> np.ndarray:
    """Element -> part by sorting centroids along one axis into `world` equal chunks (a stand-in for the
    reference's SFC / METIS partitions, whose sources are not in the checkout)."""
    c = points[axis][elements].astype(np.float64).mean(axis=0)
    order = np.argsort(c, kind="stable")
    part = np.empty(elements.shape[1], dtype=np.int32)
    bounds = np.linspace(0, order.size, world + 1).astype(np.int64)
    for p in range(world):
        part[order[bounds[p]:bounds[p + 1]]] = p
    return part


def distribute(elements: np.ndarray, n_nodes: int, elem_part: np.ndarray, rank: int, world: int) -> LocalMesh:
    """Cut one rank's LocalMesh out of a global mesh held in memory (tests, moderate sizes).

    node owner = lowest part among the elements touching it; local elements = own elements + aura
    (foreign elements touching an owned node)."""
    el = np.asarray(elements, dtype=np.int64)
    owner = np.full(n_nodes, world, dtype=np.int32)
    for v in range(el.shape[0]):
        np.minimum.at(owner, el[v], elem_part)
    own_e = np.flatnonzero(elem_part == rank)
    touches_owned = (owner[el] == rank).any(axis=0)
    aura_e = np.flatnonzero(touches_owned & (elem_part != rank))
    loc_e = np.concatenate([own_e, aura_e])
    loc_nodes = np.unique(el[:, loc_e])
    local_of_global = np.full(n_nodes, -1, dtype=np.int64)
    local_of_global[loc_nodes] = np.arange(loc_nodes.size)
    lel = local_of_global[el[:, loc_e]]
    lel, gids, own, n_owned = renumber_owned_first(lel, loc_nodes.astype(np.int64), owner[loc_nodes], rank)
    return LocalMesh(elements=np.ascontiguousarray(lel), n_owned_elements=int(own_e.size), n_owned=n_owned,
                     n_local=int(loc_nodes.size), global_ids=gids, owner=own, rank=rank, world=world)


def sshex8_box_lattice_keys(L: int, mx: int, my: int, elements: np.ndarray, n_nodes: int, z0_macro: int = 0):
    """Global lattice key (iz*nyg + iy)*nxg + ix and global iz of every node of an sshex8 box mesh whose
    macro cells are ordered z-major with x fastest (mesh.sshex8_box); z0_macro = first global macro layer."""
    from . import mesh

    ne = elements.shape[1]
    e = np.arange(ne, dtype=np.int64)
    ex, ey, ez = e % mx, (e // mx) % my, e // (mx * my)
    nxg, nyg = mx * L + 1, my * L + 1
    gid = np.empty(n_nodes, dtype=np.int64)
    izl = np.empty(n_nodes, dtype=np.int64)
    for z in range(L + 1):
        for y in range(L + 1):
            for x in range(L + 1):
                ids = elements[mesh.sshex8_lidx(L, x, y, z)]
                iz = (ez + z0_macro) * L + z
                gid[ids] = (iz * nyg + (ey * L + y)) * nxg + (ex * L + x)
                izl[ids] = iz
    return gid, izl


def sshex8_slab(level: int, mx: int, my: int, mz: int, rank: int, world: int, extent=(1.0, 1.0, 1.0)) -> LocalMesh:
    """Weak-scaling fixture: the global mesh is a stack of `world` blocks of mx*my*mz macro elements along
    z, refined to `level`; this builds rank `rank`'s LocalMesh directly (its block plus the aura layer of
    the block above), without ever materialising the global mesh. Interface planes belong to the lower rank.
    """
    from . import mesh

    L = level
    has_aura = rank < world - 1
    mz_loc = mz + (1 if has_aura else 0)
    z0 = rank * mz  # first macro layer of this rank in the global stack
    m_el, m_pts = mesh.box_hex8(mx + 1, my + 1, mz_loc + 1, 1.0, 1.0, float(mz_loc) / mz)
    m_pts[2] += np.float32(float(z0) / mz)
    if tuple(extent) != (1.0, 1.0, 1.0):  # one rank's block spans `extent` instead of the unit cube
        m_pts = (m_pts.astype(np.float64) * np.asarray(extent, dtype=np.float64)[:, None]).astype(np.float32)
    el, nn, istart, _ = mesh.sshex8_from_hex8(L, m_el, None)
    gid, izl = sshex8_box_lattice_keys(L, mx, my, el, nn, z0)
    owner = np.clip((izl + mz * L - 1) // (mz * L) - 1, 0, world - 1).astype(np.int32)
    lel, gids, own, n_owned = renumber_owned_first(el, gid, owner, rank)
    if rank > 0:
        # element order [bulk | bottom macro layer | aura]: the bottom layer reads the interface plane below (ghosts), so
        # the elements that read ghost values form ONE suffix and the distributed apply can run everything before it while
        # the halo is in flight (owned elements still come first)
        nb, no = mx * my, mx * my * mz
        lel = np.concatenate([lel[:, nb:no], lel[:, :nb], lel[:, no:]], axis=1)
    return LocalMesh(elements=np.ascontiguousarray(lel), n_owned_elements=mx * my * mz, n_owned=n_owned, n_local=nn,
                     global_ids=gids, owner=own, rank=rank, world=world,
                     extra=dict(level=L, m_elements=m_el, m_points=m_pts))


def tet4_slab(nx: int, ny: int, nz: int, rank: int, world: int, w=1.0, h=1.0, t=1.0) -> LocalMesh:
    """Rank `rank`'s share of the nx*ny*nz-node tet4 box of mesh.box_tet4 (STRONG scaling: the global mesh is fixed),
    built directly without materialising the global mesh. The box numbering x*nz + y*nz*nx + z has y slowest, so the cut
    is along y: the ny-1 cell layers are split into `world` chunks, interface node planes belong to the lower rank, the
    aura is the first cell layer of the rank above. Coordinates are slices of the global fp32 linspace, i.e. bit-identical
    to the one-piece mesh."""
    from . import mesh

    nc = ny - 1
    bounds = np.linspace(0, nc, world + 1).astype(np.int64)
    c0, c1 = int(bounds[rank]), int(bounds[rank + 1])
    has_aura = rank < world - 1
    cl = c1 - c0 + (1 if has_aura else 0)           # local cell layers
    nyl = cl + 1
    el, _ = mesh.box_tet4(nx, nyl, nz)
    gx = np.linspace(0, w, num=nx, dtype=mesh.geom_t)
    gy = np.linspace(0, h, num=ny, dtype=mesh.geom_t)[c0:c0 + nyl]
    gz = np.linspace(0, t, num=nz, dtype=mesh.geom_t)
    pts = np.stack([np.broadcast_to(gx[None, :, None], (nyl, nx, nz)).reshape(-1),
                    np.broadcast_to(gy[:, None, None], (nyl, nx, nz)).reshape(-1),
                    np.broadcast_to(gz[None, None, :], (nyl, nx, nz)).reshape(-1)]).astype(mesh.geom_t)
    plane = nx * nz
    nn = nyl * plane
    yl = np.arange(nn, dtype=np.int64) // plane       # local node layer
    gid = np.arange(nn, dtype=np.int64) + c0 * plane
    gyl = yl + c0                                     # global node layer; owner = rank whose (b_r, b_r+1] holds it
    owner = np.clip(np.searchsorted(bounds, gyl, side="left") - 1, 0, world - 1).astype(np.int32)
    n_own_e = el.shape[1]
    if has_aura:  # cells are z-major in box_tet4: move the aura layer's tets behind the owned ones (stable)
        cells_per_z = (nx - 1) * (nyl - 1)
        cell = np.arange(el.shape[1] // 6, dtype=np.int64)
        yi = (cell % cells_per_z) // (nx - 1)
        is_aura = np.repeat(yi == cl - 1, 6)          # box_tet4: element id = cell * 6 + k
        n_own_e = int(is_aura.size - is_aura.sum())
        el = np.concatenate([el[:, ~is_aura], el[:, is_aura]], axis=1)
        del is_aura
    # [owned | ghosts by owner] is a stable partition of the box numbering (y slowest): the lower ghost plane (owner
    # rank - 1) moves behind the owned block, everything else shifts down by one plane -- plain arithmetic on the ids
    # instead of the generic permutation of renumber_owned_first (6e8 tets at config-4 size)
    key = np.where(owner == rank, -1, owner).astype(np.int64)
    perm = np.argsort(key, kind="stable")
    n_owned = int((owner == rank).sum())
    lo = plane if rank > 0 else 0                     # ids [0, lo) are the lower ghost plane
    if lo:
        lel = np.where(el < lo, el + n_owned, np.where(el < lo + n_owned, el - lo, el)).astype(idx_t)
    else:
        lel = el.astype(idx_t, copy=False)
    return LocalMesh(elements=np.ascontiguousarray(lel), n_owned_elements=n_own_e, n_owned=n_owned, n_local=nn,
                     global_ids=gid[perm], owner=owner[perm].astype(np.int32), rank=rank, world=world,
                     extra=dict(points=np.ascontiguousarray(pts[:, perm])))


def coarse_space(lm: LocalMesh, corner_rows) -> tuple[LocalMesh, np.ndarray]:
    """FunctionSpace::derefine for a distributed sshex8 mesh: the coarse (macro hex8) LocalMesh whose nodes are the
    lattice corners of `lm`, numbered in the order of their fine local ids -- which keeps [owned | ghosts by owner].
    Returns (coarse mesh, fine local id of every coarse node)."""
    corners = lm.elements[np.asarray(corner_rows)]
    fine_of_coarse = np.unique(corners)
    coarse_of_fine = np.full(lm.n_local, -1, dtype=np.int64)
    coarse_of_fine[fine_of_coarse] = np.arange(fine_of_coarse.size)
    cel = coarse_of_fine[corners].astype(idx_t)
    n_owned = int((fine_of_coarse < lm.n_owned).sum())
    cm = LocalMesh(elements=np.ascontiguousarray(cel), n_owned_elements=lm.n_owned_elements, n_owned=n_owned,
                   n_local=int(fine_of_coarse.size), global_ids=lm.global_ids[fine_of_coarse], owner=lm.owner[fine_of_coarse],
                   rank=lm.rank, world=lm.world)
    return cm, fine_of_coarse


def independent_range(lm: LocalMesh):
    """(e0, e1): the longest run of local elements that read no ghost value, or None (what the overlapped distributed
    apply runs between the two halves of the halo gather)."""
    dep = (lm.elements >= lm.n_owned).any(axis=0)
    ind = np.flatnonzero(~dep)
    if ind.size == 0:
        return None
    # runs of independent elements
    cuts = np.flatnonzero(np.diff(ind) > 1)
    starts = np.concatenate([[0], cuts + 1])
    ends = np.concatenate([cuts + 1, [ind.size]])
    k = int(np.argmax(ends - starts))
    return int(ind[starts[k]]), int(ind[ends[k] - 1]) + 1


def sshex8_reachable_ghosts(lm: LocalMesh) -> np.ndarray:
    """bool [n_local - n_owned]: the ghost nodes that share a MICRO element with an owned node, i.e. the only ghost values a
    matrix-free apply can carry into an owned row (one lattice plane of the aura layer instead of all `level` of them)."""
    L = lm.extra["level"]
    N = L + 1
    el = lm.elements
    dep = np.flatnonzero((el >= lm.n_owned).any(axis=0))
    need = np.zeros(lm.n_local, dtype=bool)
    for c0 in range(0, dep.size, 4096):  # chunks: 729 x 4096 lattice ids at a time
        e = el[:, dep[c0:c0 + 4096]].reshape(N, N, N, -1)  # sshex8_lidx: z slowest, x fastest
        owned = e < lm.n_owned
        touch = np.zeros((L, L, L, e.shape[-1]), dtype=bool)  # micro elements with an owned corner
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    touch |= owned[dz:dz + L, dy:dy + L, dx:dx + L]
        reach = np.zeros_like(owned)
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    reach[dz:dz + L, dy:dy + L, dx:dx + L] |= touch
        need[e[reach]] = True
    return need[lm.n_owned:]


def halo_plan(lm: LocalMesh, all_gather_object=None, needed=None) -> dict:
    """The exchange plan of one space in the form b200fem_halo_create takes it (include/b200fem.h): per peer the
    contiguous ghost range I receive and the owned local ids I send, in the order of the peer's range.
    `all_gather_object(obj) -> [obj of every rank]` is the launcher's collective (setup only).
    needed (bool per ghost, e.g. sshex8_reachable_ghosts): exchange those ghosts only -- the plan then carries
    recv_offsets / recv_idx for b200fem_halo_set_recv_index (pass it on every rank or on none)."""
    plan = dict(n_owned=lm.n_owned, n_local=lm.n_local, send_peers=[], send_offsets=[0], send_idx=np.zeros(0, np.int32),
                recv_peers=[], recv_begin=[], recv_count=[])
    if lm.world == 1:
        return plan
    ghost_owner = lm.owner[lm.n_owned:]
    ghost_gid = lm.global_ids[lm.n_owned:]
    need = {}
    if needed is not None:
        plan["recv_offsets"], ridx = [0], []
    for p in np.unique(ghost_owner):
        sel = np.flatnonzero(ghost_owner == p)
        assert sel.size == sel[-1] - sel[0] + 1, "ghosts must be sorted by owner"
        plan["recv_peers"].append(int(p))
        plan["recv_begin"].append(lm.n_owned + int(sel[0]))
        if needed is not None:
            sel = sel[needed[sel]]
            ridx.append((lm.n_owned + sel).astype(np.int32))
            plan["recv_offsets"].append(plan["recv_offsets"][-1] + int(sel.size))
        plan["recv_count"].append(int(sel.size))
        need[int(p)] = ghost_gid[sel]
    all_need = all_gather_object(need)
    owned_gid = lm.global_ids[:lm.n_owned]
    sorter = np.argsort(owned_gid, kind="stable")
    chunks = []
    for q, nd in enumerate(all_need):
        if q == lm.rank or lm.rank not in nd:
            continue
        want = nd[lm.rank]
        pos = np.searchsorted(owned_gid, want, sorter=sorter)
        loc = sorter[np.clip(pos, 0, sorter.size - 1)]
        assert np.array_equal(owned_gid[loc], want), "a peer asked for a node this rank does not own"
        plan["send_peers"].append(q)
        plan["send_offsets"].append(plan["send_offsets"][-1] + int(loc.size))
        chunks.append(loc.astype(np.int32))
    plan["send_idx"] = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
    if needed is not None:
        plan["recv_idx"] = np.concatenate(ridx) if ridx else np.zeros(0, np.int32)
    return plan


def torch_all_gather_object(group=None):
    """all_gather_object for halo_plan under torchrun (gloo in the CPU tests, NCCL on the GPUs)."""
    import torch.distributed as dist

    def ag(obj):
        out = [None] * dist.get_world_size(group)
        dist.all_gather_object(out, obj, group=group)
        return out

    return ag


class HaloExchange:
    """`gather(x_local)`: fill the ghost block of a local vector from the owners (Exchange::gather)."""

    def __init__(self, lm: LocalMesh, block_size: int = 1, group=None, device=None, needed=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world, self.bs = lm.rank, lm.world, block_size
        self.n_owned, self.n_local = lm.n_owned, lm.n_local
        self.device = device
        self.peers, self.recv_ranges, self.send_idx, self.recv_idx = [], {}, {}, {}
        if self.world == 1:
            return
        # the same plan the C++ layer takes (b200fem_halo_create [+ b200fem_halo_set_recv_index] through host.Halo)
        self.plan = plan = halo_plan(lm, torch_all_gather_object(group), needed=needed)
        for j, (p, b, c) in enumerate(zip(plan["recv_peers"], plan["recv_begin"], plan["recv_count"])):
            self.recv_ranges[p] = (b, b + c)
            if "recv_idx" in plan:  # indexed receive: c values that go to the listed ghost nodes
                self.recv_idx[p] = torch.from_numpy(plan["recv_idx"][plan["recv_offsets"][j]:plan["recv_offsets"][j + 1]].astype(np.int64))
        for j, q in enumerate(plan["send_peers"]):
            loc = torch.from_numpy(plan["send_idx"][plan["send_offsets"][j]:plan["send_offsets"][j + 1]].copy())
            self.send_idx[q] = loc.to(device) if device is not None else loc
        self.peers = sorted(set(self.recv_ranges) | set(self.send_idx))
        self.send_buf = {}
        self.bytes_sent_per_gather = sum(int(v.numel()) for v in self.send_idx.values())

    def _send_buffer(self, q, like):
        n = self.send_idx[q].numel() * self.bs
        b = self.send_buf.get((q, like.dtype))
        if b is None or b.numel() != n:
            b = self.torch.empty(n, dtype=like.dtype, device=like.device)
            self.send_buf[(q, like.dtype)] = b
        return b

    def release(self, x_local=None):
        """Called after the local apply that consumed the ghost block has been enqueued. Nothing to do for the
        send/recv transport; the symmetric-memory transport raises "ghost block free" towards its writers."""

    def gather(self, x_local):
        """x_local: [n_local * block_size]; on return its ghost block holds the owners' values."""
        if self.world == 1 or not self.peers:
            return x_local
        torch, dist = self.torch, self.dist
        ops, staged = [], []
        for q in self.peers:
            if q in self.send_idx:
                buf = self._send_buffer(q, x_local)
                if x_local.is_cuda:
                    from . import lib

                    lib.halo_pack(self.send_idx[q], self.bs, x_local, buf)
                else:  # gloo host-logic tests only: no CUDA extension on this path
                    idx = self.send_idx[q].long()
                    buf.copy_(x_local.view(-1, self.bs)[idx].reshape(-1))
                ops.append(dist.P2POp(dist.isend, buf, q, group=self.group))
            if q in self.recv_ranges:
                a, b = self.recv_ranges[q]
                if q in self.recv_idx:
                    staged.append((q, torch.empty((b - a) * self.bs, dtype=x_local.dtype, device=x_local.device)))
                    ops.append(dist.P2POp(dist.irecv, staged[-1][1], q, group=self.group))
                else:
                    ops.append(dist.P2POp(dist.irecv, x_local[a * self.bs:b * self.bs], q, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for q, buf in staged:
            x_local.view(-1, self.bs)[self.recv_idx[q].to(x_local.device)] = buf.view(-1, self.bs)
        return x_local


class SymmetricHaloExchange(HaloExchange):
    """Halo gather without NCCL kernels: the local vector lives in NVLink-mapped symmetric memory
    (torch.distributed._symmetric_memory) and the OWNER writes the values a peer needs straight into that peer's ghost
    block -- the pack kernel (b200fem_halo_pack) with a peer pointer as destination, i.e. pack and transfer are one
    kernel of coalesced 8-byte stores over NVLink -- followed by a stream-ordered signal. Measured on 4 B200
    (tools/halo_probe.py): 0.05 ms per step against 0.21 ms for the grouped ncclSend / ncclRecv of the same 17 MB.

    Protocol per gather, all stream-ordered (signal pads are binary flags with back-pressure):
      wait "ghost block free" (channel 1) from every peer I write to   -> its previous apply has consumed the block
      push + put_signal(channel 0) to every such peer
      wait_signal(channel 0) from every peer that writes to me
    and after the local apply has been enqueued, `release()` raises "ghost block free" towards my writers.
    The vector handed to gather() must be the one returned by `local_vector()`."""

    def __init__(self, lm: LocalMesh, block_size: int = 1, dtype=None, group=None, device=None):
        super().__init__(lm, block_size, group=group, device=device)
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.dtype = dtype or torch.float64
        self._symm_mem, self._grp = symm_mem, group or dist.group.WORLD
        n = torch.tensor([self.n_local * self.bs], dtype=torch.int64, device=device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=self._grp)
        self.n_sym = int(n.item())
        # where my values land on each peer: that peer's recv range for me
        self._all_ranges = [None] * self.world
        dist.all_gather_object(self._all_ranges, self.recv_ranges, group=self._grp)
        self.vectors = {}  # data_ptr -> (tensor, handle, {peer: remote ghost-block view})
        self.buf = self.allocate()

    def allocate(self):
        """A new zero-filled symmetric vector (COLLECTIVE: every rank allocates in the same order). Any vector from
        here can be handed to gather() / release(); solvers get their work vectors this way through
        host.Blas.set_allocator."""
        torch, dist = self.torch, self.dist
        buf = self._symm_mem.empty(self.n_sym, dtype=self.dtype, device=self.device)
        hdl = self._symm_mem.rendezvous(buf, self._grp.group_name)
        buf.zero_()
        remote = {}
        for q in self.send_idx:
            a, b = self._all_ranges[q][self.rank]
            assert (b - a) == self.send_idx[q].numel()
            remote[q] = hdl.get_buffer(q, ((b - a) * self.bs,), self.dtype, a * self.bs)
        self.vectors[buf.data_ptr()] = (buf, hdl, remote)
        torch.cuda.synchronize()
        dist.barrier(group=self._grp)
        for p in self.recv_ranges:  # every ghost block starts out free
            hdl.put_signal(p, channel=1)
        return buf

    def free(self, ptr):
        self.vectors.pop(int(ptr), None)

    def local_vector(self):
        return self.buf[: self.n_local * self.bs]

    def release(self, x_local=None):
        ptr = self.buf.data_ptr() if x_local is None else (x_local if isinstance(x_local, int) else x_local.data_ptr())
        hdl = self.vectors[ptr][1]
        for p in self.recv_ranges:
            hdl.put_signal(p, channel=1)

    def gather(self, x_local):
        from . import lib

        ptr = x_local if isinstance(x_local, int) else x_local.data_ptr()
        assert ptr in self.vectors, "gather() needs a vector from local_vector() / allocate()"
        buf, hdl, remote = self.vectors[ptr]
        for q in self.send_idx:
            hdl.wait_signal(q, channel=1)
            lib.halo_pack(self.send_idx[q], self.bs, buf, remote[q])
            hdl.put_signal(q, channel=0)
        for p in self.recv_ranges:
            hdl.wait_signal(p, channel=0)
        return x_local


class OverlappedApply:
    """y_local = A x_local with the halo gather of x overlapped with the elements that read no ghost value
    (SURVEY section 8e). Element order is [ghost-dependent prefix | independent | ghost-dependent suffix] for the
    slab meshes above (bottom macro layer, bulk, aura layer); anything else falls back to gather-then-apply.

    Stream plan: the exchange (pack kernel + grouped NCCL send / recv) is enqueued FIRST on a side stream; the main
    stream zeroes the shared rows and runs the independent range on `sm_count - reserve_sms` SMs (the apply kernel's
    CTAs hold a whole SM until the launch ends, so NCCL's kernels need SMs of their own); after the exchange the two
    dependent ranges follow."""

    def __init__(self, op, lm: LocalMesh, halo: "HaloExchange", device, reserve_sms: int = 8):
        import torch

        self.torch, self.op, self.halo = torch, op, halo
        self.ne = int(lm.elements.shape[1])
        self.range = None
        if halo.world > 1 and halo.peers:
            dep = (lm.elements >= lm.n_owned).any(axis=0)
            ind = np.flatnonzero(~dep)
            if ind.size and not dep[ind[0]:ind[-1] + 1].any():
                self.range = (int(ind[0]), int(ind[-1]) + 1)
        self.reserve_sms = reserve_sms
        self.comm = torch.cuda.Stream(device=device, priority=-1) if self.range else None

    def apply_overwrite(self, x_local, y_local):
        op, halo = self.op, self.halo
        if self.range is None:
            halo.gather(x_local)
            out = op.apply_overwrite(x_local, y_local)
            halo.release(x_local)
            return out
        torch = self.torch
        main = torch.cuda.current_stream()
        self.comm.wait_stream(main)  # x_local is ready
        with torch.cuda.stream(self.comm):
            halo.gather(x_local)
        e0, e1 = self.range
        op.set_option("reserve_sms", self.reserve_sms)
        op.apply_range(x_local, y_local, e0, e1, 1)
        op.set_option("reserve_sms", 0)
        main.wait_stream(self.comm)
        if e0 > 0:
            op.apply_range(x_local, y_local, 0, e0, 2)
        if e1 < self.ne:
            op.apply_range(x_local, y_local, e1, self.ne, 2)
        halo.release(x_local)
        return y_local

    def launches(self):
        return 1 + len(self.halo.send_idx) + (self.op.launches(True) + 2 if self.range else self.op.launches(True))


class ParallelMatrixFreeOperator:
    """Operator<T> over the owned dofs: y = A x with halo gather (rows() = owned dofs,
    row_allocation_size() = local dofs), same contract as sfem::ParallelMatrixFreeOperator."""

    def __init__(self, local_op, exchange: HaloExchange, x_local, y_local):
        self.op, self.ex = local_op, exchange
        self.x_local, self.y_local = x_local, y_local
        self.owned_dofs = exchange.n_owned * exchange.bs
        self.local_dofs = exchange.n_local * exchange.bs

    def rows(self):
        return self.owned_dofs

    def row_allocation_size(self):
        return self.local_dofs

    def apply(self, x_owned, y_owned):
        """y_owned = A x_owned (the reference zeroes its scratch output first, so this IS its result)."""
        xl, yl = self.x_local, self.y_local
        xl[: self.owned_dofs].copy_(x_owned)
        self.ex.gather(xl)
        self.op.apply_overwrite(xl, yl)
        self.ex.release(xl)
        y_owned.copy_(yl[: self.owned_dofs])
        return y_owned

    def apply_local(self):
        """Same, for callers that keep their vectors at local size (x in self.x_local, y in self.y_local)."""
        self.ex.gather(self.x_local)
        self.op.apply_overwrite(self.x_local, self.y_local)
        self.ex.release(self.x_local)

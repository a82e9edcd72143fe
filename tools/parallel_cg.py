#!/usr/bin/env python
"""Distributed CG (ParallelConjugateGradient + ParallelMatrixFreeOperator of the reference,
algebra/parallel/sfem_parallel_cg.hpp:257-400, frontend/sfem_ParallelMatrixFreeOperator.cpp:78-105) on N GPUs:

    torchrun --nproc-per-node N tools/parallel_cg.py [--macro 64] [--level 8] [--max-it 200]

One rank per GPU; the global mesh is the z-stack of bench.py (sfem_b200.parallel.sshex8_slab). The solver is the
C++ host layer's ConjugateGradient over the OWNED dofs, with
  * Operator<T>::apply = copy owned -> local, NCCL halo gather of the ghost block, local y = A x through the C ABI
    (with the Dirichlet rows of the local Function), copy local -> owned;
  * every dot product = local fused kernel + allreduce(SUM) (DeviceBLAS::allreduce_sum).
With --check (small sizes) rank 0 also solves the same global problem on one GPU and compares iteration counts and
residual histories -- the reference's own parity criterion for its distributed path
(frontend/tests/mpi/sfem_ParallelLaplacianApplyTest.cpp:133-141), extended to the solver."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist

    from sfem_b200 import host, lib, mesh, parallel

    ap = argparse.ArgumentParser()
    ap.add_argument("--macro", type=int, default=64)
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--max-it", type=int, default=200)
    ap.add_argument("--rtol", type=float, default=1e-30)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--nccl-halo", action="store_true", help="grouped ncclSend/ncclRecv instead of the symmetric-memory push")
    a = ap.parse_args()

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dt, L = torch.float64, a.level

    lm = parallel.sshex8_slab(L, a.macro, a.macro, a.macro, rank, world)
    m_el, m_pts = lm.extra["m_elements"], lm.extra["m_points"]
    fff = lib.fff_fill(lib.HEX8, m_el, m_pts)
    op = lib.laplacian(lib.SSHEX8, lm.elements, fff, lm.n_local, dt, level=L)
    f = host.Function(dt, lm.n_local, 1)
    f.add_operator(op)
    # u = -1 / +1 on the x-faces (present on every rank of a z-stack), zero rhs
    f.add_dirichlet(mesh.sshex8_face_nodes_macro(L, lm.elements, m_el, m_pts, 0, 0.0), 0, -1.0)
    f.add_dirichlet(mesh.sshex8_face_nodes_macro(L, lm.elements, m_el, m_pts, 0, 1.0), 0, 1.0)
    n_owned, n_local = lm.n_owned, lm.n_local
    symmetric = world > 1 and not a.nccl_halo
    if symmetric:  # owner-push over NVLink-mapped memory (parallel.SymmetricHaloExchange)
        halo = parallel.SymmetricHaloExchange(lm, 1, dtype=dt, device=dev)
        xl = halo.local_vector()
    else:
        halo = parallel.HaloExchange(lm, 1, device=dev)
        xl = torch.zeros(n_local, dtype=dt, device=dev)
    yl = torch.zeros(n_local, dtype=dt, device=dev)
    raw = lib.raw()
    scal = torch.zeros(1, dtype=torch.float64, device=dev)

    hraw = host.raw()

    def apply_in_place(x_addr, y_addr):
        # symmetric path: every solver vector is a symmetric vector with room for the ghost block (allocator hook
        # below), so the exchange and the apply work on the solver's own vectors: no copies
        halo.gather(x_addr)
        rc = hraw.b200host_apply_overwrite(f.ptr, x_addr, y_addr)
        halo.release(x_addr)
        return rc

    def apply_owned(x_addr, y_addr):  # y_owned = A x_owned through the local scratch vectors
        raw.b200fem_memcpy_d2d(xl.data_ptr(), x_addr, n_owned * 8, None)
        halo.gather(xl)
        host.apply_overwrite(f, xl, yl)
        if symmetric:
            halo.release()
        raw.b200fem_memcpy_d2d(y_addr, yl.data_ptr(), n_owned * 8, None)
        return 0

    def apply_owned_acc(x_addr, y_addr):
        raise RuntimeError("the solvers use the overwriting form")

    def allreduce(v):
        if world == 1:
            return v
        scal[0] = v
        dist.all_reduce(scal)
        return float(scal.item())

    A = host.callback_op(dt, n_owned, n_owned, apply_owned_acc, overwrite=apply_in_place if symmetric else apply_owned)
    blas = host.Blas(dt, allreduce=allreduce if world > 1 else None)
    if symmetric:
        blas.set_allocator(lambda nbytes: halo.allocate().data_ptr(), halo.free)
        x, x_warm = halo.allocate(), halo.allocate()
    else:
        x, x_warm = torch.zeros(n_owned, dtype=dt, device=dev), torch.zeros(n_owned, dtype=dt, device=dev)
    b = torch.zeros(n_owned, dtype=dt, device=dev)
    f.apply_constraints(xl)  # sets the Dirichlet values on the local vector; owned part is what the solver sees
    x[:n_owned].copy_(xl[:n_owned])
    x_warm[:n_owned].copy_(xl[:n_owned])
    b.copy_(xl[:n_owned])
    xl.zero_()
    cg = host.cg(blas, A, None, rtol=a.rtol, atol=1e-30, max_it=a.max_it)
    host.apply(cg, b, x_warm)  # warm-up (allocates the work vectors)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = host.apply(cg, b, x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tot = torch.tensor([float(n_owned), ms], dtype=torch.float64, device=dev)
    if world > 1:
        t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        n_all = torch.tensor([float(n_owned)], dtype=torch.float64, device=dev)
        dist.all_reduce(n_all)
        tot = torch.stack([n_all[0], t_ms[0]])
    hist = host.residual_history(cg)
    it = host.iterations(cg)
    out = None
    if rank == 0:
        dofs, ms = float(tot[0]), float(tot[1])
        out = {"name": f"distributed CG Poisson, sshex8 L{L}, {a.macro}^3 macro per GPU, z-stack x{world}", "n_gpus": world,
               "dofs_total": int(dofs), "iterations": it, "rc": rc, "ms_total": ms, "ms_per_iteration": ms / max(it, 1),
               "gdof_s_per_iteration": dofs / (ms / max(it, 1) * 1e-3) / 1e9, "residual_0": float(hist[0]),
               "residual_end": float(hist[-1]), "halo_values_sent_rank0": halo.bytes_sent_per_gather if world > 1 else 0}
    if a.check:
        # the same global problem on one GPU (rank 0 only): iteration count and residual history must agree
        if rank == 0:
            # the slab stack spans z in [0, world] (every rank's block is a unit cube): same geometry
            m = mesh.sshex8_box(L, a.macro, a.macro, a.macro * world, 1.0, 1.0, float(world))
            fff1 = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
            op1 = lib.laplacian(lib.SSHEX8, m["elements"], fff1, m["n_nodes"], dt, level=L)
            f1 = host.Function(dt, m["n_nodes"], 1)
            f1.add_operator(op1)
            f1.add_dirichlet(mesh.sshex8_face_nodes(L, m["elements"], m["m_points"], 0, 0.0), 0, -1.0)
            f1.add_dirichlet(mesh.sshex8_face_nodes(L, m["elements"], m["m_points"], 0, 1.0), 0, 1.0)
            x1 = torch.zeros(m["n_nodes"], dtype=dt, device=dev)
            f1.apply_constraints(x1)
            b1 = x1.clone()
            cg1 = host.cg(host.Blas(dt), f1, None, rtol=a.rtol, atol=1e-30, max_it=a.max_it)
            rc1 = host.apply(cg1, b1, x1)
            h1 = host.residual_history(cg1)
            out["check"] = {"serial_iterations": host.iterations(cg1), "serial_rc": rc1,
                            "max_rel_residual_history_diff": float(np.max(np.abs(h1 - hist) / np.maximum(np.abs(h1), 1e-300)))
                            if len(h1) == len(hist) else None}
            assert host.iterations(cg1) == it and rc1 == rc, (host.iterations(cg1), it)
            # the macro coordinates (fp32, geom_t) of the stacked blocks and of the one-piece box round differently, so
            # the two FFF sets agree to ~1e-7 only: iteration counts must be identical, residual histories close
            assert out["check"]["max_rel_residual_history_diff"] < 1e-4, out["check"]
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

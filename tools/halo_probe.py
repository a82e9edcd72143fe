#!/usr/bin/env python
"""Where does a distributed apply step go? torchrun --nproc-per-node N tools/halo_probe.py : times, per rank and as the
max over ranks, the halo gather alone, the local apply alone and the two together (not a bench line)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from sfem_b200 import lib, parallel

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
macro = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lm = parallel.sshex8_slab(8, macro, macro, macro, rank, world)
fff = lib.fff_fill(lib.HEX8, lm.extra["m_elements"], lm.extra["m_points"])
op = lib.laplacian(lib.SSHEX8, lm.elements, fff, lm.n_local, torch.float64, level=8)
x = torch.rand(lm.n_local, dtype=torch.float64, device=dev)
y = torch.zeros_like(x)
halo = parallel.HaloExchange(lm, 1, device=dev)


def timed(fn, steps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    return [round(float(v), 4) for v in allt]


def both():
    halo.gather(x)
    op.apply_overwrite(x, y)


def pack_only():
    for q in halo.peers:
        if q in halo.send_idx:
            lib.halo_pack(halo.send_idx[q], 1, x, halo._send_buffer(q, x))


def exchange_only():
    ops = []
    for q in halo.peers:
        if q in halo.send_idx:
            ops.append(dist.P2POp(dist.isend, halo._send_buffer(q, x), q))
        if q in halo.recv_ranges:
            a, b = halo.recv_ranges[q]
            ops.append(dist.P2POp(dist.irecv, x[a:b], q))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def exchange_and_apply():
    exchange_only()
    op.apply_overwrite(x, y)


sym = parallel.SymmetricHaloExchange(lm, 1, dtype=torch.float64, device=dev)
xs = sym.local_vector()
xs.copy_(x)
halo.gather(x)          # reference: NCCL gather
xs[lm.n_owned:] = -7.0  # poison the ghost block, then fill it through the symmetric push
torch.cuda.synchronize(); dist.barrier()
sym.gather(xs)
torch.cuda.synchronize(); dist.barrier()
sym_equal = bool(torch.equal(xs, x))
sym.release()


def sym_step():
    sym.gather(xs)
    op.apply_overwrite(xs, y)
    sym.release()


def sym_gather_only():
    sym.gather(xs)
    sym.release()


res = {"symmetric_equals_nccl": sym_equal, "symmetric_gather_only_ms": timed(sym_gather_only),
       "symmetric_gather_then_apply_ms": timed(sym_step), "pack_only_ms": timed(pack_only), "exchange_only_ms": timed(exchange_only), "exchange_then_apply_ms": timed(exchange_and_apply),
       "gather_only_ms": timed(lambda: halo.gather(x)), "apply_only_ms": timed(lambda: op.apply_overwrite(x, y)),
       "gather_then_apply_ms": timed(both), "values_sent": halo.bytes_sent_per_gather,
       "peers": {int(q): int(v.numel()) for q, v in halo.send_idx.items()}, "elements": int(lm.elements.shape[1])}
allres = [None] * world
dist.all_gather_object(allres, {k: res[k] for k in ("values_sent", "peers", "elements")})
if rank == 0:
    res["per_rank"] = allres
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()

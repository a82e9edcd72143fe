#!/usr/bin/env python
"""bench.py -- headline benchmark of the matrix-free apply hot path (BASELINE.json metric).

A "step" is one matrix-free Laplacian apply over the sshex8 mesh of config C3:
level-8 semi-structured hex8 on a 64^3 macro mesh per GPU, 135 005 697 dofs, fp64, synthetic unit cube.
What is timed is  y = A x  -- the pair `zeros(y); Op::apply(x, y)` every solver of the reference issues
(algebra/sfem_cg.hpp:226-230, sfem_Multigrid.hpp:271), fused into one C-ABI call
(b200fem_op_apply_overwrite). The strict accumulate form  y += A x  is timed beside it.

With N > 1 ranks (torchrun, one process per GPU) the global mesh is an N-long stack of such blocks along z
(weak scaling). Every rank holds its block plus the aura layer of the block above
(sfem_b200/parallel.py, the reference's owned/ghost/aura scheme) and every step is one apply of the C++ host
layer's ParallelMatrixFreeOperator (halo gather of x: owner-push over NVLink-mapped memory, else grouped ncclSend /
ncclRecv; then the local kernel). `value` counts the dofs all ranks own. The solver rows of BASELINE.json ride in
`solves`: config 4 (tet4 CG, strong scaling over the ranks) and config 5 (sshex8 multigrid elasticity, weak scaling:
87 x 87 x 11 macro elements per GPU = 1.03e9 dofs on 8), both through tools/dist_solve.py.

Prints ONE JSON line (DESIGN.md "Measurement" explains every field). `--impl reference` times the
reference's own OpenMP kernel (oracle/_ref, else the oracle port) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "sshex8_laplacian_apply_throughput"
UNIT = "GDOF/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--macro", type=int, default=64, help="macro elements per direction per rank")
    ap.add_argument("--level", type=int, default=8)
    ap.add_argument("--real", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: halo gather in front of the apply (default: the gather on the halo's own stream while the apply kernel runs, its "
                         "ghost-dependent element groups wait for the ghost block inside the kernel)")
    ap.add_argument("--full-halo", action="store_true", help="N > 1: exchange every ghost node of the aura elements, not only those that can reach an owned row")
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: grouped ncclSend/ncclRecv instead of the owner-push transport")
    ap.add_argument("--c4-side", type=int, default=465, help="nodes per side of the config-4 tet4 cube (strong scaling)")
    ap.add_argument("--c5-macro", type=int, default=87, help="config 5: macro elements per x / y side")
    ap.add_argument("--c5-layers", type=int, default=11, help="config 5: macro layers per GPU (z-stack)")
    ap.add_argument("--no-solves", action="store_true", help="skip the CG / multigrid solve times (tools/solver_bench.py)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs 0 / 1 / 2-fp32 (tools/extras_bench.py)")
    ap.add_argument("--extras", action="store_true", help="also time the other configs (C1, C2, fp32, vector kernels)")
    return ap.parse_args()


def workload_name(a):
    return f"C3: sshex8 Laplacian apply, level {a.level}, {a.macro}^3 macro elements per GPU, {a.real}"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.3)  # let the first samples land before the timed region starts
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own OpenMP kernel on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(a, elements, fff, nn, steps, warmup):
    """affine_sshex8_laplacian_stencil_apply_fff from oracle/_ref (the reference's sources, its Release
    flags) -- or the oracle port when the reference build is absent -- with all host threads."""
    from oracle.ref import Oracle, available

    cores = os.cpu_count() or 1
    if "TORCHELASTIC_RUN_ID" in os.environ or "OMP_NUM_THREADS" not in os.environ:
        # torchrun exports OMP_NUM_THREADS=1 for every rank; only rank 0 runs this leg and it gets all host cores
        os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.setdefault("OMP_PROC_BIND", "true")
    kind = "reference" if available("reference", a.real, fast=True) else "port"
    isa = "avx2+fma"
    orc = None
    if kind == "reference" and "avx512f" in cpu_flags():
        try:  # the reference builds with -march=native; the box's ISA is not known at build time, so two builds travel
            orc = Oracle(kind, a.real, fast="avx512")
            isa = "x86-64-v4 (avx512)"
        except Exception:
            orc = None
    if orc is None:
        orc = Oracle(kind, a.real, fast=(kind == "reference"))
    try:  # libgomp may have been initialised (by numpy / torch) before the variable was set
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
    except Exception:
        pass
    u = np.random.default_rng(42).uniform(-1, 1, nn).astype(orc.rt)
    y = np.zeros(nn, dtype=orc.rt)
    for _ in range(warmup):
        orc.sshex8_laplacian_apply(a.level, elements, fff, u, out=y)
    t0 = time.perf_counter()
    for _ in range(steps):
        y[:] = 0  # the reference's callers zero the output before every apply
        orc.sshex8_laplacian_apply(a.level, elements, fff, u, out=y)
    dt = (time.perf_counter() - t0) / steps
    return {"value": nn / dt / 1e9, "unit": UNIT, "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": kind,
            "sample": f"the full workload ({nn} dofs), {steps} x (zero y + apply) after {warmup} warm-up, wall clock, "
                      f"OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']}",
            "cpu_model": cpu_model(), "isa": isa if kind == "reference" else "-O2 port",
            "host_GB_s": nn * (2 * orc.rt().itemsize) / dt / 1e9 + elements.size * 4 / dt / 1e9,
            "ms_per_step": dt * 1e3}


def host_memory_gb():
    """memory this process may still take: MemAvailable capped by the cgroup limit"""
    avail = 0.0
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = float(ln.split()[1]) / 1e6
    except Exception:
        return 0.0
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(path).read().strip()
            if v != "max":
                used = 0.0
                try:
                    used = float(open(path.replace("memory.max", "memory.current").replace("limit_in_bytes", "usage_in_bytes")).read())
                except Exception:
                    pass
                avail = min(avail, (float(v) - used) / 1e9)
        except Exception:
            pass
    return avail


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_flags():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("flags"):
                return set(ln.split(":", 1)[1].split())
    except Exception:
        pass
    return set()


def workload_config(a):
    """`config` of the JSON line: the same dict in both arms (what differs between them is said in `config_details`)"""
    return {"workload": workload_name(a),
            "semantics": "y = A x: the reference's zeros(y) + Op::apply(x, y) pair (one call, b200fem_op_apply_overwrite, in the b200 arm)",
            "l2_policy": "x + y + connectivity = 2.2 GB per apply, larger than the 126 MB L2; no flush"}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.ref import Oracle, available
    from sfem_b200 import mesh

    m = mesh.sshex8_box(a.level, a.macro, a.macro, a.macro, with_points=False)
    kind = "reference" if available("reference", "f64") else "port"
    fff = Oracle(kind, "f64").hex8_fff_fill(m["m_elements"], m["m_points"])
    steps, warm = max(1, a.steps), max(0, a.warmup)  # the caller's counts: one step = 0.4 s on 16 cores
    cb = cpu_reference_run(a, m["elements"], fff, m["n_nodes"], steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.real, "data": "synthetic",
        # `config` is the same dict in both arms (the driver compares them); what differs is said beside it
        "config": workload_config(a),
        "config_details": {"arm": "the reference's OpenMP kernel (oracle/_ref) on the host cores of rank 0, one rank's block"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model", "isa", "host_GB_s")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# b200 arm
# ------------------------------------------------------------------------------------------------------
def time_steps(fn, steps, warmup, torch, barrier=None):
    """CUDA events on the current stream (the one every C-ABI call is launched on), sync on both sides."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return ev0.elapsed_time(ev1) / steps  # ms per step


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this rank's host threads (and with them the first-touch placement of its pinned e2e buffers) to the NUMA node
    its GPU hangs off: with every rank's buffers on node 0 the 8-GPU e2e copies all cross one socket."""
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_b200(a):
    import torch

    from sfem_b200 import host, lib, mesh, parallel
    from tools.dist_solve import Dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)  # setup plumbing (unique id, plan objects) and the driver's NCCL check
    d = Dist(torch)  # the C++ communicator every timed collective goes through

    dt = torch.float64 if a.real == "f64" else torch.float32
    ndt = np.float64 if a.real == "f64" else np.float32
    t0 = time.time()
    if world == 1:
        m = mesh.sshex8_box(a.level, a.macro, a.macro, a.macro, with_points=False)
        elements, n_local, n_owned = m["elements"], m["n_nodes"], m["n_nodes"]
        m_el, m_pts = m["m_elements"], m["m_points"]
        lm = None
    else:
        lm = parallel.sshex8_slab(a.level, a.macro, a.macro, a.macro, rank, world)
        elements, n_local, n_owned = lm.elements, lm.n_local, lm.n_owned
        m_el, m_pts = lm.extra["m_elements"], lm.extra["m_points"]
    t_mesh = time.time() - t0
    fff = lib.fff_fill(lib.HEX8, m_el, m_pts)  # macro-element FFF from the product's own geometry kernel
    t0 = time.time()
    op = lib.laplacian(lib.SSHEX8, elements, fff, n_local, dt, level=a.level)
    t_plan = time.time() - t0
    assert op.get_option("mode") == 1 and op.get_option("interior_exclusive") == 1

    u_host = np.random.default_rng(42 + rank).uniform(-1, 1, n_local).astype(ndt)
    x = torch.from_numpy(u_host).to(dev)
    y = torch.zeros_like(x)
    halo, halo_kind, A = None, None, None
    if lm is not None:
        # ParallelMatrixFreeOperator of the C++ host layer: halo gather of x (in place, into its ghost block) + local apply
        f = host.Function(dt, n_local, 1)
        f.add_operator(op)
        # only the ghosts that share a micro element with an owned node travel (one lattice plane of the aura layer instead of
        # all `level` of them: the others cannot reach an owned row); --full-halo: every node of the aura elements, like the
        # reference's Exchange
        needed = None if a.full_halo else parallel.sshex8_reachable_ghosts(lm)
        halo = host.Halo(d.comm, dt, 1, parallel.halo_plan(lm, d.all_gather_object, needed=needed),
                         transport=lib.HALO_NCCL if a.nccl_halo else lib.HALO_AUTO)
        # the elements that read no ghost value run between the two halves of the gather (--no-overlap: gather, then apply)
        ind = None if a.no_overlap else parallel.independent_range(lm)
        A = host.parallel_op(f, halo, n_owned, n_owned, independent_range=ind)
        halo_values = halo.values_sent
        halo_kind = {"push": "owner-push: the owner's pack kernel stores into the peer's receive block over NVLink (cudaIpc-mapped), "
                             "acquire / release epoch flags, a second kernel waits and copies into the ghost block",
                     "nccl": "pack kernel + grouped ncclSend / ncclRecv into the ghost block"}[halo.transport]

    def step():
        if A is not None:
            rc = host.apply_overwrite(A, x, y)
            assert rc == 0
        else:
            op.apply_overwrite(x, y)

    barrier = d.comm.barrier if world > 1 else None
    max_over_ranks = d.comm.max

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = max_over_ranks(time_steps(step, a.steps, a.warmup, torch, barrier))
    clocks = sampler.stop() if rank == 0 else None

    total_owned = int(d.comm.sum(n_owned))

    # the dominant kernel alone (no halo): overwrite form = [memset of the skeleton ranges + stream kernel]
    ms_kernel = time_steps(lambda: op.apply_overwrite(x, y), a.steps, 2, torch)
    ms_accumulate = time_steps(lambda: op.apply(x, y), max(3, a.steps // 2), 2, torch)

    # e2e: HOST buffers through the C ABI, H2D of x and D2H of y inside the timed region
    e2e = None
    if not a.no_e2e:
        xh = torch.from_numpy(u_host).pin_memory()
        yh = torch.empty_like(xh).pin_memory()

        # Double-buffered pipeline over three streams: the H2D copy of step k+1 and the D2H copy of step k-1 run
        # while step k computes (PCIe is full duplex; one step moves 1.08 GB each way against 1 ms of kernel time).
        # Every step still copies its whole input from pinned host memory and its whole result back.
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        xb = [x, torch.empty_like(x)]
        yb = [y, torch.empty_like(y)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        state = {"k": 0}

        def e2e_step():
            k = state["k"]
            b = k % 2
            state["k"] = k + 1
            with torch.cuda.stream(s_in):
                if k >= 2:
                    s_in.wait_event(ev_comp[b])  # the apply of step k-2 has consumed this x buffer
                xb[b].copy_(xh, non_blocking=True)
                ev_in[b].record(s_in)
            main.wait_event(ev_in[b])
            if k >= 2:
                main.wait_event(ev_out[b])  # the D2H of step k-2 has drained this y buffer
            if A is not None:
                host.apply_overwrite(A, xb[b], yb[b])
            else:
                op.apply_overwrite(xb[b], yb[b])
            ev_comp[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_comp[b])
                yh.copy_(yb[b], non_blocking=True)
                ev_out[b].record(s_out)

        def e2e_run(nsteps):
            state["k"] = 0
            for _ in range(nsteps):
                e2e_step()
            main.wait_stream(s_out)  # the timed region ends when the last result is on the host

        k = max(4, min(a.steps, 8))
        e2e_run(2)  # warm-up
        torch.cuda.synchronize()
        if barrier:
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(k)
        e1.record()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1) / k)
        # the result that reached the host is the apply of the host input
        chk = float((yh[:n_owned].to(dev) - yb[(k - 1) % 2][:n_owned]).abs().max())
        assert chk == 0.0, chk
        e2e = {"value": total_owned / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(n_local * x.element_size()), "d2h_bytes_per_step": int(n_local * x.element_size()),
               "ms_per_step": ms_e2e, "steps": k,
               "note": "every step: pinned host x -> device, y = A x through the C ABI, device y -> pinned host; double-buffered "
                       "over three streams so that the copies of neighbouring steps overlap (PCIe-bound: "
                       f"{2 * n_local * x.element_size() / (ms_e2e * 1e-3) / 1e9:.1f} GB/s in + out)"}
        del xb, yb

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg, streamed, op_launches = op.algorithmic_bytes, op.streamed_bytes(True), op.launches(True)
    halo_launches = halo.launches_per_gather if halo is not None else 0
    # the gather on the halo's own stream, the ghost-dependent element groups wait inside the apply kernel (same launch count)
    overlapped = bool(A is not None and getattr(A, "overlapped", False))
    n_elements = int(elements.shape[1])

    # the solver rows of BASELINE.json (configs 4 and 5) on all ranks, through the C++ distributed layer
    solves = None
    if not a.no_solves:
        from tools import dist_solve, solver_bench

        del op, x, y, A, halo
        if lm is not None:
            del f
        torch.cuda.empty_cache()
        solves = []

        def attempt(fn, *args, **kw):
            try:
                solves.append(fn(*args, **kw))
            except Exception as ex:  # report, do not hide the headline number
                solves.append({"name": getattr(fn, "__name__", "solve"), "error": repr(ex)})
            torch.cuda.empty_cache()

        if world == 1:
            attempt(solver_bench.cg_poisson, torch, macro=a.macro, peak_gbs=peak)
        # config 4, STRONG scaling: the 465^3-node cube cut into `world` slabs. One GPU holds it too (6e8 tets: 24 GB of
        # connectivity + FFF) when the host has the memory to generate it; else its 1/8 share (233^3) is reported
        side = a.c4_side
        if world == 1 and host_memory_gb() < 160:
            side = (a.c4_side + 1) // 2
        attempt(dist_solve.cg_poisson_tet4, d, n=side, peak_gbs=peak)
        # config 5, WEAK scaling: 87 x 87 x 11 macro elements per GPU (1.03e9 dofs on 8 GPUs)
        attempt(dist_solve.mg_elasticity, d, macro=a.c5_macro, layers=a.c5_layers)

    if rank != 0:
        d.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    traffic, traffic_source = None, None  # dram__bytes_read + dram__bytes_write of the dominant kernel (ncu --set full)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if a.real == "f64" and a.macro == 64 and a.level == 8 and world == 1:
            traffic = tr["dram_bytes_per_launch"]
            traffic_source = "committed capture: " + tr.get("source", "profiles/traffic.json")
    except Exception:
        pass
    achieved = alg / (ms_kernel * 1e-3) / 1e9
    launches_per_step = op_launches + halo_launches
    line = {
        "metric": METRIC, "value": total_owned / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.real, "data": "synthetic",
        "config": workload_config(a),
        "config_details": {
            "dofs_owned_total": total_owned, "dofs_local_rank0": n_local,
            "macro_elements_rank0": n_elements,
            "connectivity": "27 affine boxes per macro element (432 B instead of 2916 B of int32 indices)",
            "parallelism": (f"z-slab x{world}, owned/ghost/aura; every step = ParallelMatrixFreeOperator::apply of the C++ host "
                            f"layer: halo gather of x ({halo_kind}; "
                            + ("every ghost node of the aura elements" if a.full_halo else "the ghost nodes that share a micro element with an owned "
                               "node only, the others cannot reach an owned row") + f": {halo_values} values sent by rank 0), "
                            + ("on the halo's own stream while the apply kernel runs; its ghost-dependent element groups, ordered last, "
                               "wait for the ghost block inside the kernel" if overlapped else "then the local kernel")
                            + (f"; rank 0 bound to NUMA node {numa_node}" if numa_node is not None else ""))
            if world > 1 else "single GPU",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_source, "algorithmic_bytes": alg, "streamed_bytes": streamed,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
            "kernel": ("sshex8_laplacian_ws2_kernel" if (a.real == "f64" and a.level == 8) else "sshex8_laplacian_stream_kernel")
            + " (+ the memsets of the skeleton ranges that belong to the call)",
            "kernel_ms": ms_kernel, "frac_of_8TBs_nominal": achieved / 8000.0,
        },
        "apply_accumulate": {"value": n_local / (ms_accumulate * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_accumulate,
                             "note": "strict y += A x on rank 0's local mesh (b200fem_op_apply)"},
        "gpu_launches": launches_per_step * a.steps,
        "clocks": clocks,
        "setup_s": {"mesh": t_mesh, "plan": t_plan},
    }
    if e2e:
        line["e2e"] = e2e
    if not a.no_cpu_baseline and world == 1:
        try:
            cb = cpu_reference_run(a, elements, fff, n_local, 20, 2)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "cpu_model", "isa", "host_GB_s")}
        except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
    if solves is not None:
        line["solves"] = solves
    if not a.no_other_configs and world == 1:
        # the other single-GPU configurations of BASELINE.json beside the headline (config 0: tet4 Laplacian on the
        # ~1 M-element cube; config 1: hex8 linear elasticity 128^3; config 2 in fp32), same timing rules
        from tools.extras_bench import run_extras

        try:
            line["other_configs"] = run_extras(torch, which=("c1s", "c2", "c3f32"))
        except Exception as ex:
            line["other_configs"] = [{"error": repr(ex)}]
    if a.extras:
        from tools.extras_bench import run_extras

        line["extras"] = run_extras(torch)
    print(json.dumps(line), flush=True)
    d.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()

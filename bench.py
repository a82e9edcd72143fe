#!/usr/bin/env python
"""bench.py -- headline benchmark of the matrix-free apply hot path (BASELINE.json metric).

A "step" is one matrix-free Laplacian apply over the sshex8 mesh of config C3:
level-8 semi-structured hex8 on a 64^3 macro mesh per GPU, 135 005 697 dofs, fp64, synthetic unit cube.
What is timed is  y = A x  -- the pair `zeros(y); Op::apply(x, y)` every solver of the reference issues
(algebra/sfem_cg.hpp:226-230, sfem_Multigrid.hpp:271), fused into one C-ABI call
(b200fem_op_apply_overwrite). The strict accumulate form  y += A x  is timed beside it.

With N > 1 ranks (torchrun, one process per GPU) the global mesh is an N-long stack of such blocks along z
(weak scaling). Every rank holds its block plus the aura layer of the block above
(sfem_b200/parallel.py, the reference's owned/ghost/aura scheme) and every step starts with the NCCL halo
gather of x. `value` counts the dofs all ranks own.

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
    ap.add_argument("--nccl-halo", action="store_true", help="N > 1: grouped ncclSend/ncclRecv instead of the symmetric-memory push")
    ap.add_argument("--overlap", action="store_true",
                    help="N > 1: overlap the halo gather with the ghost-independent elements (b200fem_op_apply_range). Off by "
                         "default: measured on 2 GPUs it is slower (1.19 vs 1.10 ms per step) -- three launches of the "
                         "persistent kernel and 8 reserved SMs cost more than the 0.12 ms exchange they hide")
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
            "ms_per_step": dt * 1e3}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle.ref import Oracle, available
    from sfem_b200 import mesh

    m = mesh.sshex8_box(a.level, a.macro, a.macro, a.macro, with_points=False)
    kind = "reference" if available("reference", "f64") else "port"
    fff = Oracle(kind, "f64").hex8_fff_fill(m["m_elements"], m["m_points"])
    steps = max(1, min(a.steps, 20))
    warm = max(1, min(a.warmup, 3))
    cb = cpu_reference_run(a, m["elements"], fff, m["n_nodes"], steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.real, "data": "synthetic",
        "config": {"workload": workload_name(a) + " -- reference OpenMP kernel on the host cores (one rank's block)"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
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


def run_b200(a):
    import torch

    from sfem_b200 import lib, mesh, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

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
    halo, halo_kind = None, None
    if lm is not None:
        if not a.nccl_halo:
            try:  # owner-push over NVLink-mapped symmetric memory (sfem_b200.parallel.SymmetricHaloExchange)
                halo = parallel.SymmetricHaloExchange(lm, 1, dtype=dt, device=dev)
                halo_kind = "owner-push into the peers' ghost blocks over symmetric (NVLink-mapped) memory + stream-ordered signals"
            except Exception as ex:  # no P2P mapping on this box: grouped ncclSend / ncclRecv
                halo_kind = f"grouped ncclSend/ncclRecv (symmetric memory unavailable: {type(ex).__name__})"
        if halo is None:
            halo = parallel.HaloExchange(lm, 1, device=dev)
            halo_kind = halo_kind or "grouped ncclSend/ncclRecv"
    if isinstance(halo, parallel.SymmetricHaloExchange):
        x = halo.local_vector()
        x.copy_(torch.from_numpy(u_host))
    else:
        x = torch.from_numpy(u_host).to(dev)
    y = torch.zeros_like(x)
    symmetric = isinstance(halo, parallel.SymmetricHaloExchange)
    overlapped = parallel.OverlappedApply(op, lm, halo, dev) if (halo is not None and a.overlap) else None

    def step():
        if overlapped is not None:
            overlapped.apply_overwrite(x, y)
            if symmetric:
                halo.release()
            return
        if halo is not None:
            halo.gather(x)
        op.apply_overwrite(x, y)
        if symmetric:
            halo.release()  # my ghost block may be overwritten again once this apply has run

    if overlapped is not None and overlapped.range:
        # one-off check: the overlapped schedule computes what gather-then-apply computes (differences = RED order)
        y_ref = torch.zeros_like(y)
        halo.gather(x)
        op.apply_overwrite(x, y_ref)
        if symmetric:
            halo.release()
        overlapped.apply_overwrite(x, y)
        if symmetric:
            halo.release()
        rel = float((y[:n_owned] - y_ref[:n_owned]).norm() / y_ref[:n_owned].norm())
        assert rel < 1e-12, f"overlapped apply differs from gather-then-apply: {rel}"
        del y_ref

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = max_over_ranks(time_steps(step, a.steps, a.warmup, torch, barrier))
    clocks = sampler.stop() if rank == 0 else None

    total_owned = n_owned
    if dist is not None:
        t = torch.tensor([n_owned], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        total_owned = int(t.item())

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
        if symmetric:
            xb = [x, halo.allocate()[: x.numel()]]
        else:
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
            if halo is not None:
                halo.gather(xb[b])
            op.apply_overwrite(xb[b], yb[b])
            if symmetric:
                halo.release(xb[b])
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
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(k)
        e1.record()
        torch.cuda.synchronize()
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

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None  # dram__bytes_read + dram__bytes_write of the dominant kernel, from the committed ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if a.real == "f64" and a.macro == 64 and a.level == 8:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    alg = op.algorithmic_bytes
    achieved = alg / (ms_kernel * 1e-3) / 1e9
    launches_per_step = op.launches(True) + (len(halo.send_idx) if halo is not None else 0)
    line = {
        "metric": METRIC, "value": total_owned / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.real, "data": "synthetic",
        "config": {
            "workload": workload_name(a), "dofs_owned_total": total_owned, "dofs_local_rank0": n_local,
            "macro_elements_rank0": int(elements.shape[1]),
            "semantics": "y = A x: the reference's zeros(y) + Op::apply(x, y) pair as one call (b200fem_op_apply_overwrite)",
            "l2_policy": "x + y + connectivity = 2.2 GB per apply, larger than the 126 MB L2; no flush",
            "connectivity": "27 affine boxes per macro element (432 B instead of 2916 B of int32 indices)",
            "parallelism": (f"z-slab x{world}, owned/ghost/aura, halo gather of x per step: {halo_kind}"
                            + (", overlapped with the ghost-independent elements" if overlapped is not None and overlapped.range else ""))
            if world > 1 else "single GPU",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "algorithmic_bytes": alg, "streamed_bytes": op.streamed_bytes(True),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s",
            "kernel": ("sshex8_laplacian_ws_kernel" if (a.real == "f64" and a.level == 8) else "sshex8_laplacian_stream_kernel")
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
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
    if not a.no_solves and world == 1:
        from tools.solver_bench import run_solves

        del op, x, y
        torch.cuda.empty_cache()
        line["solves"] = run_solves(torch, cg_macro=a.macro, mg_macro=max(4, a.macro // 2), peak_gbs=peak)
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

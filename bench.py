#!/usr/bin/env python
"""bench.py -- headline benchmark of the matrix-free apply hot path (BASELINE.json metric).

A "step" is one matrix-free Laplacian apply  y += A x  (the reference's Op::apply contract,
frontend/ops/sfem_Op.hpp:214) over the whole sshex8 mesh of config C3:
level-8 semi-structured hex8 on a 64^3 macro mesh, 135 005 697 dofs, fp64, synthetic unit cube.
Per rank the mesh is one such block (weak scaling); with N > 1 ranks the global mesh is an N-long
stack of blocks along z and every apply ends with the interface exchange (NCCL).

Prints ONE JSON line (see DESIGN.md "Measurement" for every field). `--impl reference` times the
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
    ap.add_argument("--extras", action="store_true", help="also time the other configs (C1, C2, fp32, overwrite)")
    return ap.parse_args()


def workload_name(a):
    return f"sshex8 Laplacian apply, level {a.level}, {a.macro}^3 macro elements per GPU, {a.real}"


def build_mesh(a):
    from sfem_b200 import mesh

    t0 = time.time()
    m = mesh.sshex8_box(a.level, a.macro, a.macro, a.macro, with_points=False)
    m["t_mesh"] = time.time() - t0
    return m


def macro_fff(m):
    """FFF of the macro elements (unit cube box: computed with the product's own geometry kernel)."""
    from sfem_b200 import lib

    return lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(a, m, fff, steps, warmup):
    """The reference's own OpenMP kernel (oracle/_ref `_fast` build = its Release flags) on the host."""
    from oracle.ref import Oracle, available

    kind = "reference" if available("reference", a.real, fast=True) else "port"
    orc = Oracle(kind, a.real, fast=(kind == "reference"))
    nn = m["n_nodes"]
    u = np.random.default_rng(42).uniform(-1, 1, nn).astype(orc.rt)
    y = np.zeros(nn, dtype=orc.rt)
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    os.environ.setdefault("OMP_PROC_BIND", "true")
    for _ in range(warmup):
        orc.sshex8_laplacian_apply(a.level, m["elements"], fff, u, out=y)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.sshex8_laplacian_apply(a.level, m["elements"], fff, u, out=y)
    dt = (time.perf_counter() - t0) / steps
    return {"value": nn / dt / 1e9, "unit": UNIT, "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": kind,
            "sample": f"full workload ({nn} dofs), {steps} applies after {warmup} warm-up, wall clock",
            "ms_per_step": dt * 1e3}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m = build_mesh(a)
    from oracle.ref import Oracle, available

    kind = "reference" if available("reference", "f64") else "port"
    fff = Oracle(kind, "f64").hex8_fff_fill(m["m_elements"], m["m_points"])
    steps = max(1, min(a.steps, 10))
    warm = max(1, min(a.warmup, 2))
    cb = cpu_reference_run(a, m, fff, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.real, "data": "synthetic",
        "config": {"workload": workload_name(a) + " (reference OpenMP kernel on host cores, one rank's block)"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_steps(fn, steps, warmup, torch, barrier=None):
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

    from sfem_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dt = torch.float64 if a.real == "f64" else torch.float32
    m = build_mesh(a)
    fff = macro_fff(m)
    nn = m["n_nodes"]
    t0 = time.time()
    op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, dt, level=a.level)
    t_plan = time.time() - t0

    u_host = np.random.default_rng(42 + rank).uniform(-1, 1, nn).astype(np.float64 if a.real == "f64" else np.float32)
    x = torch.from_numpy(u_host).cuda()
    y = torch.zeros_like(x)

    halo = None
    if world > 1:
        from sfem_b200.parallel import SlabHalo

        halo = SlabHalo(m, a.level, a.macro, rank, world, dt)

    def step():
        op.apply(x, y)
        if halo is not None:
            halo.exchange_add(y)

    def barrier():
        if dist is not None:
            dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = time_steps(step, a.steps, a.warmup, torch, barrier)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # kernel-only time of the dominant kernel (no halo), for the roofline object
    ms_kernel = time_steps(lambda: op.apply(x, y), a.steps, 1, torch) if halo is not None else ms
    ms_overwrite = time_steps(lambda: op.apply_overwrite(x, y), max(3, a.steps // 2), 1, torch)

    # e2e: host buffers, H2D of x and D2H of y inside the timed region, through the C ABI
    e2e = None
    if not a.no_e2e:
        xh = torch.from_numpy(u_host).pin_memory()
        yh = torch.empty_like(xh).pin_memory()

        def e2e_step():
            x.copy_(xh, non_blocking=True)
            y.zero_()
            op.apply(x, y)
            if halo is not None:
                halo.exchange_add(y)
            yh.copy_(y, non_blocking=True)

        k = max(2, min(a.steps, 5))
        ms_e2e = time_steps(e2e_step, k, 1, torch, barrier)
        if dist is not None:
            t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        e2e = {"value": nn * world / (ms_e2e * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nn * x.element_size()),
               "d2h_bytes_per_step": int(nn * x.element_size()), "ms_per_step": ms_e2e, "steps": k}

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
    alg = op.algorithmic_bytes
    achieved = alg / (ms_kernel * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": nn * world / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": a.real, "data": "synthetic",
        "config": {
            "workload": workload_name(a), "dofs_per_gpu": nn, "macro_elements_per_gpu": int(m["elements"].shape[1]),
            "semantics": "y += A x (accumulate)", "l2_policy": "inputs (2.9 GB/apply) larger than L2, no flush",
            "connectivity": "boxes" if op.get_option("mode") == 1 else "explicit",
            "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "algorithmic_bytes": alg, "streamed_bytes": op.streamed_bytes(False),
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
            "kernel_ms": ms_kernel,
        },
        "apply_overwrite": {"value": nn / (ms_overwrite * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_overwrite,
                            "note": "y = A x (fused zero + apply, what the device CG uses)"},
        "gpu_launches": op.launches(False) * a.steps,
        "clocks": clocks,
        "setup_s": {"mesh": m["t_mesh"], "plan": t_plan},
    }
    if e2e:
        line["e2e"] = e2e
    if not a.no_cpu_baseline and world == 1:
        try:
            cb = cpu_reference_run(a, m, fff, 5, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(ex)}
    if a.extras:
        from sfem_b200.extras_bench import run_extras

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

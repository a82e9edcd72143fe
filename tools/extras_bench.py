"""Secondary measurements for bench.py --extras: the other BASELINE.json configs that fit one GPU and the
vector kernels, each as apply throughput + fraction of the measured HBM copy bandwidth. Not bench lines."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def _time(fn, torch, steps=20, warmup=3, flush=None):
    """ms per call; with `flush` (a big tensor) L2 is overwritten between calls and each call timed alone."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    tot = 0.0
    for _ in range(steps):
        flush.add_(1.0)  # 512 MB read + write: evicts the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def _entry(name, rows, alg_bytes, ms, extra=None):
    d = {"name": name, "rows": int(rows), "ms": ms, "gdof_s": rows / ms / 1e6, "alg_GB_s": alg_bytes / ms / 1e6,
         "frac_of_measured_hbm": alg_bytes / ms / 1e6 / _peak()}
    if extra:
        d.update(extra)
    return d


def run_extras(torch, which=("c1", "c2", "hex8lap", "c3f32", "c5le", "em", "transfer", "blas")):
    from sfem_b200 import lib, mesh

    out = []
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
    SC = {"patch": lib.SCATTER_PATCH, "atomic": lib.SCATTER_ATOMIC}
    if "c1" in which or "c1s" in which:  # tet4 Laplacian fp64, 56^3 nodes (BASELINE config 1) and a 100x larger cube
        for n in ((56, 260) if "c1" in which else (56,)):
            el0, pts0 = mesh.box_tet4(n, n, n)
            for order in ("generator order", "SFC order"):
                el, pts = (el0, pts0) if order == "generator order" else mesh.sfc_reorder(el0, pts0)[:2]
                fff = lib.fff_fill(lib.TET4, el, pts)
                nn = pts.shape[1]
                op = lib.laplacian(lib.TET4, el, fff, nn, torch.float64)
                x = torch.rand(nn, dtype=torch.float64, device="cuda")
                y = torch.zeros_like(x)
                for scatter in ("patch", "atomic"):
                    op.set_option("scatter", SC[scatter])
                    extra = {"scatter": scatter, "order": order, "patch_avg_unique_nodes": op.get_option("patch_avg_unique_nodes")}
                    ms = _time(lambda: op.apply_overwrite(x, y), torch, flush=flush if n == 56 else None)
                    out.append(_entry(f"C1 tet4 Laplacian f64 {n}^3 nodes ({scatter}, {order}, y = A x" + (", L2 flushed" if n == 56 else "") + ")",
                                      nn, op.algorithmic_bytes, ms, extra))
                    if n == 56:
                        ms = _time(lambda: op.apply_overwrite(x, y), torch)
                        out.append(_entry(f"C1 tet4 Laplacian f64 {n}^3 nodes ({scatter}, {order}, y = A x, L2-resident)", nn,
                                          op.algorithmic_bytes, ms, extra))
                del op
    if "c2" in which or "hex8lap" in which:
        el0, pts0 = mesh.box_hex8(129, 129, 129)
        for order in ("generator order", "SFC order"):
            el, pts = (el0, pts0) if order == "generator order" else mesh.sfc_reorder(el0, pts0)[:2]
            nn = pts.shape[1]
            if "c2" in which:  # hex8 linear elasticity fp64, 128^3 elements (BASELINE config 2)
                adj, det = lib.adjugate_and_det_fill(lib.HEX8, el, pts)
                for rt, dt in (("f64", torch.float64), ("f32", torch.float32)):
                    op = lib.linear_elasticity(lib.HEX8, el, adj, det, 1.0, 1.0, nn, dt)
                    x = torch.rand(nn * 3, dtype=dt, device="cuda")
                    y = torch.zeros_like(x)
                    for scatter in ("patch", "atomic"):
                        op.set_option("scatter", SC[scatter])
                        ms = _time(lambda: op.apply_overwrite(x, y), torch)
                        out.append(_entry(f"C2 hex8 linear elasticity {rt} 128^3 ({scatter}, {order}, y = A x)", nn * 3,
                                          op.algorithmic_bytes, ms, {"scatter": scatter, "order": order}))
                    del op
            if "hex8lap" in which:
                fff = lib.fff_fill(lib.HEX8, el, pts)
                op = lib.laplacian(lib.HEX8, el, fff, nn, torch.float64)
                x = torch.rand(nn, dtype=torch.float64, device="cuda")
                y = torch.zeros_like(x)
                for scatter in ("patch", "atomic"):
                    op.set_option("scatter", SC[scatter])
                    ms = _time(lambda: op.apply_overwrite(x, y), torch)
                    out.append(_entry(f"hex8 Laplacian f64 128^3 ({scatter}, {order}, y = A x)", nn, op.algorithmic_bytes, ms,
                                      {"scatter": scatter, "order": order}))
                del op
    if "c3f32" in which:
        m = mesh.sshex8_box(8, 64, 64, 64, with_points=False)
        fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
        nn = m["n_nodes"]
        op = lib.laplacian(lib.SSHEX8, m["elements"], fff, nn, torch.float32, level=8)
        x = torch.rand(nn, dtype=torch.float32, device="cuda")
        y = torch.zeros_like(x)
        ms = _time(lambda: op.apply_overwrite(x, y), torch)
        out.append(_entry("C3 sshex8 Laplacian L8 f32 64^3 macro (y = A x)", nn, op.algorithmic_bytes, ms))
        del op, m
    if "c5le" in which or "transfer" in which:
        # one GPU's share of BASELINE config 5: sshex8 L8 linear elasticity, 44^3 macro elements (~131 M dofs)
        mm = 44
        m = mesh.sshex8_box(8, mm, mm, mm, with_points=False)
        nn = m["n_nodes"]
        if "c5le" in which:
            adj, det = lib.adjugate_and_det_fill(lib.HEX8, m["m_elements"], m["m_points"])
            for rt, dt in (("f64", torch.float64), ("f32", torch.float32)):
                op = lib.linear_elasticity(lib.SSHEX8, m["elements"], adj, det, 1.0, 1.0, nn, dt, level=8)
                x = torch.rand(nn * 3, dtype=dt, device="cuda")
                y = torch.zeros_like(x)
                ms = _time(lambda: op.apply_overwrite(x, y), torch, steps=10)
                flops = 2.0 * 444 * 512 * m["elements"].shape[1]
                out.append(_entry(f"C5 shard sshex8 linear elasticity L8 {rt} {mm}^3 macro (y = A x)", nn * 3,
                                  op.algorithmic_bytes, ms, {"fp_ops_TFLOPs": flops / ms / 1e9}))
                del op, x, y
        if "transfer" in which:
            tr = lib.Transfer(8, 1, m["elements"], nn, 1, 8, m["elements"], nn)
            for vs in (1, 3):
                f = torch.rand(nn * vs, dtype=torch.float64, device="cuda")
                c = torch.zeros(nn * vs, dtype=torch.float64, device="cuda")
                ms = _time(lambda: tr.restrict(f, c, vs), torch, steps=10)
                out.append(_entry(f"sshex8 hierarchical restriction L8 -> 1, vec_size {vs}, f64", nn * vs,
                                  nn * vs * 8.0 + nn * 2.0 + m["elements"].size * 4.0, ms))
                ms = _time(lambda: tr.prolongate(c, f, vs), torch, steps=10)
                out.append(_entry(f"sshex8 hierarchical prolongation 1 -> L8, vec_size {vs}, f64", nn * vs,
                                  nn * vs * 8.0 + m["elements"].size * 4.0, ms))
                del f, c
            del tr
        del m
    if "em" in which:
        # stored-element-matrix semi-structured operators (ss:em:Laplacian / ss:em:LinearElasticity): the reference's fastest
        # published GPU formulation (24.6 - 38 GDOF/s on its hardware, notebook/SEMISTRUCTURED.md). B_alg counts the stored
        # matrix as the geometry stream: (8 bs)^2 reals per macro element.
        m = mesh.sshex8_box(8, 64, 64, 64, with_points=False)
        nn = m["n_nodes"]
        fff = lib.fff_fill(lib.HEX8, m["m_elements"], m["m_points"])
        for rt, dt in (("f64", torch.float64), ("f32", torch.float32)):
            M = lib.sshex8_laplacian_element_matrix(8, fff, dt)
            op = lib.sshex8_em(8, m["elements"], M, 1, nn, dt)
            x = torch.rand(nn, dtype=dt, device="cuda")
            y = torch.zeros_like(x)
            ms = _time(lambda: op.apply_overwrite(x, y), torch, steps=10)
            out.append(_entry(f"ss:em:Laplacian L8 {rt} 64^3 macro (stored 8x8 per macro element, y = A x)", nn, op.algorithmic_bytes, ms))
            del op, x, y
        del m
        mm = 44
        m = mesh.sshex8_box(8, mm, mm, mm, with_points=False)
        nn = m["n_nodes"]
        adj, det = lib.adjugate_and_det_fill(lib.HEX8, m["m_elements"], m["m_points"])
        for rt, dt in (("f64", torch.float64), ("f32", torch.float32)):
            K = lib.sshex8_linear_elasticity_element_matrix(8, adj, det, 1.0, 1.0, dt)
            op = lib.sshex8_em(8, m["elements"], K, 3, nn, dt)
            x = torch.rand(nn * 3, dtype=dt, device="cuda")
            y = torch.zeros_like(x)
            ms = _time(lambda: op.apply_overwrite(x, y), torch, steps=5)
            out.append(_entry(f"ss:em:LinearElasticity L8 {rt} {mm}^3 macro (stored 24x24 per macro element, y = A x)", nn * 3,
                              op.algorithmic_bytes, ms, {"fp_ops_TFLOPs": 2.0 * 576 * 512 * m["elements"].shape[1] / ms / 1e9}))
            del op, x, y
        del m
    if "p2" in which:  # section 8 f3 / f4: 10-node tetrahedra and the packed (uint16) Laplacian, 80^3-corner cube in SFC order
        el4, pts4 = mesh.sfc_reorder(*mesh.box_tet4(80, 80, 80))[:2]
        el10, pts10 = mesh.tet4_to_tet10(el4, pts4)
        nn4, nn10 = pts4.shape[1], pts10.shape[1]
        fff = lib.fff_fill(lib.TET4, el4, pts4)
        adj, det = lib.adjugate_and_det_fill(lib.TET4, el4, pts4)
        for fam, et in (("tet10", lib.TET10), ("macro_tet4", lib.MACRO_TET4)):
            for kind in ("laplacian", "linear elasticity"):
                bs = 1 if kind == "laplacian" else 3
                op = (lib.laplacian(et, el10, fff, nn10, torch.float64) if bs == 1 else
                      lib.linear_elasticity(et, el10, adj, det, 1.0, 1.0, nn10, torch.float64))
                x = torch.rand(nn10 * bs, dtype=torch.float64, device="cuda")
                y = torch.zeros_like(x)
                for scatter in ("patch", "atomic"):
                    op.set_option("scatter", SC[scatter])
                    ms = _time(lambda: op.apply_overwrite(x, y), torch)
                    out.append(_entry(f"{fam} {kind} f64 {el10.shape[1]} elements ({scatter}, SFC order, y = A x)", nn10 * bs,
                                      op.algorithmic_bytes, ms, {"scatter": scatter, "patch_avg_unique_nodes": op.get_option("patch_avg_unique_nodes")}))
                del op, x, y
        for fam, et, el, nn, epp in (("tet4", lib.TET4, el4, nn4, 1024), ("tet10", lib.TET10, el10, nn10, 512)):
            pm = lib.PackedMesh(el, nn, epp)
            op = lib.packed_laplacian(et, pm, fff, torch.float64)  # FFF depends on the element only, not on the node numbering
            ref = lib.laplacian(et, el, fff, nn, torch.float64)
            x = torch.rand(nn, dtype=torch.float64, device="cuda")
            y = torch.zeros_like(x)
            # packed bytes: 2 B per corner + ghost ids, against 4 B per corner
            alg = ref.algorithmic_bytes
            for direct in (0, 1):
                op.set_option("packed_kernel", direct)
                ms = _time(lambda: op.apply_overwrite(x, y), torch)
                out.append(_entry(f"packed {fam} laplacian f64 {el.shape[1]} elements ({'direct' if direct else 'staged'} kernel, {epp} per pack, y = A x)",
                                  nn, alg, ms, {"streamed_bytes": op.streamed_bytes(True), "max_nodes_per_pack": pm.max_nodes_per_pack,
                                                "ghost_entries": pm.n_ghost_entries, "alg_bytes_are": "the unpacked operator's (4 B indices)"}))
            ms = _time(lambda: ref.apply_overwrite(x, y), torch)
            out.append(_entry(f"unpacked {fam} laplacian f64 {el.shape[1]} elements (patch kernel, same mesh, y = A x)", nn, alg, ms))
            del op, ref, x, y, pm
    if "blas" in which:
        from sfem_b200.lib import blas

        n = 135_005_697
        a, b, c, d = (torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(4))
        for name, words, fn in (
            ("dot", 2, lambda: blas.dot(a, b)),
            ("axpby", 3, lambda: blas.axpby(0.5, a, 0.25, b)),
            ("zaxpby", 3, lambda: blas.zaxpby(0.5, a, 0.25, b, c)),
            ("cg_update_x_r_rtr", 6, lambda: blas.cg_update_x_r_rtr(1e-3, a, b, c, d)),
            ("cg_update_p", 3, lambda: blas.cg_update_p(0.5, a, b)),
            ("cheb3_update", 6, lambda: blas.cheb3_update(1e-3, 0.5, a, b, c, d)),
        ):
            ms = _time(fn, torch, steps=10)
            out.append(_entry(f"blas {name} f64 n=135M", n, words * 8.0 * n, ms))
    return out


if __name__ == "__main__":
    import sys

    import torch

    sys.path.insert(0, ROOT)
    which = tuple(sys.argv[1].split(",")) if len(sys.argv) > 1 else None
    for e in (run_extras(torch, which) if which else run_extras(torch)):
        print(json.dumps(e), flush=True)

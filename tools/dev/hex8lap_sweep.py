import json, sys
sys.path.insert(0, "/root/repo")
import torch
from sfem_b200 import lib, mesh
from tools.extras_bench import _entry, _time
el0, pts0 = mesh.box_hex8(129, 129, 129)
for order in ("generator order", "SFC order"):
    el, pts = (el0, pts0) if order == "generator order" else mesh.sfc_reorder(el0, pts0)[:2]
    nn = pts.shape[1]
    fff = lib.fff_fill(lib.HEX8, el, pts)
    op = lib.laplacian(lib.HEX8, el, fff, nn, torch.float64)
    x = torch.rand(nn, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x); ref = torch.zeros_like(x)
    op.set_option("scatter", lib.SCATTER_ATOMIC); op.apply_overwrite(x, ref)
    for sc, name in ((lib.SCATTER_ATOMIC, "atomic"), (lib.SCATTER_PATCH, "patch")):
        op.set_option("scatter", sc)
        ms = _time(lambda: op.apply_overwrite(x, y), torch)
        e = _entry(f"hex8 Laplacian f64 128^3 ({name}, {order})", nn, op.algorithmic_bytes, ms)
        e["rel_vs_atomic"] = float((y - ref).norm() / ref.norm())
        print(json.dumps(e), flush=True)

#!/usr/bin/env python
"""Per-role (mover / compute) instruction mix and stall samples of a warp-specialised kernel from an ncu report:
   python tools/dev/ncu_roles.py report.ncu-rep [warp_steps]   (roles are split at the USETMAXREG instructions)"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 786432.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
print("kernel", d.get("Kernel Name", "")[:90])
for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum"]:
    if k in d: print(" ", k, d[k])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
marks = [k for k, r in enumerate(data) if "USETMAXREG" in r[ix["Source"]]]
def agg(lo, hi, name):
    ex, sm, st = collections.Counter(), collections.Counter(), collections.Counter()
    S = E = 0
    for r in data[lo:hi]:
        toks = r[ix["Source"]].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "SYNCS")) else op.split(".")[0]
        s, e = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        ex[op] += e; sm[op] += s; S += s; E += e
        for h in stalls: st[h] += int(r[ix[h]])
    print(f"== {name}: samples {S}, executed {E} = {E / steps:.1f} per warp step, SASS lines {hi - lo}")
    print("   ", ", ".join(f"{op} {e / steps:.1f}" for op, e in ex.most_common(26)))
    print("    stalls:", ", ".join(f"{h[6:]} {c}" for h, c in st.most_common(10)))
agg(marks[0], marks[1], "mover")
agg(marks[1], len(data), "compute")

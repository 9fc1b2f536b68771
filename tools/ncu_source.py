#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an ncu report (source page; no GPU needed).
usage: ncu_source.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
b = blocks[0]
h = b["hdr"]
ci = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_math", "stall_mio",
                              "stall_lg", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_no_inst")}
tot = sum(int(r[ci["# Samples"]]) for r in b["rows"])
print(f"# {b['name']}: {len(b['rows'])} SASS instructions, {tot} warp-stall samples")
agg = {k: sum(int(r[ci[k]]) for r in b["rows"]) for k in ci if k.startswith("stall")}
print("# samples by reason:", {k: f"{v / tot:.3f}" for k, v in sorted(agg.items(), key=lambda x: -x[1])})
print(f"{'idx':>5s} {'samples':>8s} {'share':>6s} {'execs':>9s} {'long_sb':>7s} {'short':>6s} {'bar':>5s} {'wait':>5s} {'math':>5s} {'shWave':>8s} {'ideal':>8s}  instruction")
order = sorted(range(len(b["rows"])), key=lambda i: -int(b["rows"][i][ci["# Samples"]]))[:top]
for i in sorted(order):
    r = b["rows"][i]
    print(f"{i:5d} {r[ci['# Samples']]:>8s} {int(r[ci['# Samples']]) / tot:6.3f} {r[ci['Instructions Executed']]:>9s} {r[ci['stall_long_sb']]:>7s} {r[ci['stall_short_sb']]:>6s} "
          f"{r[ci['stall_barrier']]:>5s} {r[ci['stall_wait']]:>5s} {r[ci['stall_math']]:>5s} {r[ci['L1 Wavefronts Shared']]:>8s} {r[ci['L1 Wavefronts Shared Ideal']]:>8s}  {r[ci['Source']].strip()}")

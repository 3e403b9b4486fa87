#!/usr/bin/env python
"""Opcode histogram of one kernel of an .ncu-rep, weighted by EXECUTED warp instructions (SASS page of the report):
tools/ncu_opcodes.py <rep> <kernel substring> [out.csv] -> opcode, static count, executed warp instructions, share, lanes"""
import collections, csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2].replace(" ", "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif row and row[0] == "Address":
        cur["hdr"] = row
    elif row and cur is not None and "hdr" in cur:
        cur["rows"].append(row)
blk = [b for b in blocks if pat in b["name"].replace("(int)", "").replace("(bool)", "").replace(" ", "")][0]
h = {n: i for i, n in enumerate(blk["hdr"])}
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r in blk["rows"]:
    op = r[h["Source"]].split()
    op = [t for t in op if not t.startswith("@")][0].split(".")[0].rstrip(";")
    wi, ti = int(r[h["Instructions Executed"]]), int(r[h["Thread Instructions Executed"]])
    agg[op][0] += 1; agg[op][1] += wi; agg[op][2] += ti
    tot += wi
rows = sorted(agg.items(), key=lambda x: -x[1][1])
out = [["opcode", "static_instructions", "executed_warp_instructions", "share_pct", "active_lanes"]]
for op, (n, wi, ti) in rows:
    out.append([op, n, wi, "%.2f" % (100.0 * wi / tot), "%.1f" % (ti / max(1, wi))])
if len(sys.argv) > 3:
    with open(sys.argv[3], "w", newline="") as f:
        csv.writer(f).writerows(out)
print(blk["name"], "static", len(blk["rows"]), "executed", tot)
for r in out[:26]:
    print("  %-12s %6s %14s %7s %6s" % tuple(r))

#!/usr/bin/env python
"""Join an ncu SASS-level source page with nvdisasm line info of the SAME build:
tools/ncu_lines.py <rep> <so> <kernel substring> [top] -> per source line: stall samples, warp instructions executed,
lane efficiency.  Also aggregates by the outermost inlined-at frame (function-level view)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, so, pat = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
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
blk = [b for b in blocks if pat.replace(" ", "") in b["name"].replace("(int)", "").replace("(bool)", "").replace(" ", "")][0]
h = {n: i for i, n in enumerate(blk["hdr"])}
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
lines = subprocess.run(["nvdisasm", "--print-line-info-inline", cub], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.")]
def mangled(pat):
    # "k_env_step<1, 0, 1>" -> "k_env_stepILi1ELb0ELb1E"; "k_movegen<false>" / "k_movegen<0>" -> "k_movegenILb0E"
    m = re.match(r"(\w+)<(.*)>$", pat.replace(" ", ""))
    if not m:
        return pat
    name, args = m.group(1), m.group(2).split(",")
    if name == "k_env_step":  # <int MODE, int TILE, bool SELFPLAY>
        args = (args + ["0", "0"])[:3]
        return name + "ILi%sELi%sELb%dE" % (args[0], args[1], int(args[2] not in ("0", "false")))
    return name + "I" + "".join("Lb%dE" % int(a not in ("0", "false")) for a in args)
mang = mangled(pat)
locs = []
for si, st in enumerate(starts):
    if mang not in lines[st]:
        continue
    en = starts[si + 1] if si + 1 < len(starts) else len(lines)
    frames = []
    fresh = True
    for l in lines[st:en]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            if fresh:
                frames, fresh = [], False
            frames.append((m.group(1).split("/")[-1], int(m.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
            locs.append(tuple(frames))
            fresh = True
    break
rows = blk["rows"]
assert len(rows) == len(locs), (len(rows), len(locs))
by_line, by_outer = collections.defaultdict(lambda: [0, 0, 0]), collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r, frames in zip(rows, locs):
    s, wi, ti = int(r[h["# Samples"]]), int(r[h["Instructions Executed"]]), int(r[h["Thread Instructions Executed"]])
    loc = frames[0] if frames else ("?", 0)
    # function-level view: the frame one below the outermost two (kernel line -> env_step_one line -> ...)
    outer = frames[-2] if len(frames) >= 2 else loc
    if len(frames) >= 3 and frames[-2][0] == "env_core.cuh":
        outer = frames[-3] if frames[-3][0] == "env_core.cuh" else frames[-2]
    for dct, k in ((by_line, loc), (by_outer, outer)):
        dct[k][0] += s; dct[k][1] += wi; dct[k][2] += ti
    tot[0] += s; tot[1] += wi; tot[2] += ti
print("total samples %d warp-inst %d lane-eff %.1f" % (tot[0], tot[1], tot[2] / max(1, tot[1])))
for title, dct in (("by innermost line", by_line), ("by env_core-level frame", by_outer)):
    print("==", title)
    for k, (s, wi, ti) in sorted(dct.items(), key=lambda x: -x[1][0])[:top]:
        print("  %-24s %5d  samples %6d (%4.1f%%)  warp-inst %9d (%4.1f%%)  lanes %.1f" % (k[0], k[1], s, 100.0 * s / tot[0], wi, 100.0 * wi / tot[1], ti / max(1, wi)))
csv_out = os.environ.get("NCU_LINES_CSV")  # NCU_LINES_CSV=<path>: the per-line table (top `top` lines by warp instructions) as CSV
if csv_out:
    src_cache = {}
    def src_line(f, ln):
        for cand in ("gym_chess_b200/csrc/" + f, f):
            if os.path.exists(cand):
                if cand not in src_cache:
                    src_cache[cand] = open(cand).read().split("\n")
                return src_cache[cand][ln - 1].strip()[:100] if 0 < ln <= len(src_cache[cand]) else ""
        return ""
    with open(csv_out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["file", "line", "stall_samples", "samples_pct", "warp_instructions", "warp_instructions_pct", "active_lanes", "source"])
        for k, (sm, wi, ti) in sorted(by_line.items(), key=lambda x: -x[1][1])[:top]:
            w.writerow([k[0], k[1], sm, "%.2f" % (100.0 * sm / tot[0]), wi, "%.2f" % (100.0 * wi / tot[1]), "%.1f" % (ti / max(1, wi)), src_line(k[0], k[1])])
if len(sys.argv) > 5:  # annotated listing: file:lo-hi
    f, rng = sys.argv[5].split(":")
    lo, hi = map(int, rng.split("-"))
    path = [p for p in ("gym_chess_b200/csrc/" + f, f) if os.path.exists(p)][0]
    src = open(path).read().split("\n")
    for ln in range(lo, hi + 1):
        s, wi, ti = by_line.get((f, ln), (0, 0, 0))
        print("%5d %6d %9d %5.1f | %s" % (ln, s, wi, ti / max(1, wi), src[ln - 1][:110]))

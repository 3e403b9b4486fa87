#!/usr/bin/env python
"""Function-level view of an ncu source page: tools/ncu_funcs.py <rep> <so> <kernel> -> per source FUNCTION (every inlined-at
frame mapped to the function whose body holds the line) inclusive and exclusive warp instructions, samples, lane efficiency."""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile
sys.argv_saved = sys.argv
rep, so, pat = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
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
m = re.match(r"(\w+)<(.*)>$", pat.replace(" ", ""))
args = m.group(2).split(",")
mang = m.group(1) + ("ILi%sELi%sELb%dE" % (args[0], args[1], int(args[2] not in ("0", "false"))) if m.group(1) == "k_env_step" else "")
locs = []
for si, st in enumerate(starts):
    if mang not in lines[st]:
        continue
    en = starts[si + 1] if si + 1 < len(starts) else len(lines)
    frames, fresh = [], True
    for l in lines[st:en]:
        mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if mm:
            if fresh: frames, fresh = [], False
            frames.append((mm.group(1).split("/")[-1], int(mm.group(2))))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
            locs.append(tuple(frames)); fresh = True
    break
rows = blk["rows"]
assert len(rows) == len(locs), (len(rows), len(locs))
def funcs(path):
    out = []
    for i, l in enumerate(open(path), 1):
        mm = re.match(r"^(?:GCB_HD|__device__ __forceinline__|__global__|static __device__|template <[^>]*>\s*GCB_HD)\s+.*?(\w+)\s*\(", l)
        if mm and mm.group(1) not in ("defined", "__launch_bounds__"): out.append((i, mm.group(1)))
        mm = re.match(r"^\s+(?:GCB_HD|__device__ __forceinline__)\s+.*?(\w+)\s*\(", l)   # member functions
        if mm: out.append((i, mm.group(1)))
    return out
F = {f: funcs("gym_chess_b200/csrc/" + f) for f in ("chess_core.cuh", "env_core.cuh", "gcb_kernels.cu")}
def fn(frame):
    f, n = frame
    if f not in F: return f
    ls = [x[0] for x in F[f]]
    i = bisect.bisect_right(ls, n) - 1
    return F[f][i][1] if i >= 0 else f + ":?"
incl, excl = collections.defaultdict(lambda: [0, 0, 0]), collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for r, frames in zip(rows, locs):
    s, wi, ti = int(r[h["# Samples"]]), int(r[h["Instructions Executed"]]), int(r[h["Thread Instructions Executed"]])
    names = [fn(fr) for fr in frames] or ["?"]
    for k in set(names):
        incl[k][0] += s; incl[k][1] += wi; incl[k][2] += ti
    excl[names[0]][0] += s; excl[names[0]][1] += wi; excl[names[0]][2] += ti
    tot[0] += s; tot[1] += wi; tot[2] += ti
print("total samples %d warp-inst %d lanes %.1f" % (tot[0], tot[1], tot[2] / max(1, tot[1])))
for title, dct in (("inclusive", incl), ("exclusive (innermost frame)", excl)):
    print("==", title)
    for k, (s, wi, ti) in sorted(dct.items(), key=lambda x: -x[1][1])[:40]:
        print("  %-28s samples %6d (%4.1f%%)  warp-inst %10d (%4.1f%%)  lanes %.1f" % (k, s, 100.0 * s / tot[0], wi, 100.0 * wi / tot[1], ti / max(1, wi)))

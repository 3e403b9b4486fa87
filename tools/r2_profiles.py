#!/usr/bin/env python
"""Rebuild everything under profiles/ that is derived from the captures of tools/r2_record.sh (run here, after the
.ncu-rep files and the launch list came back in gpurun_out/):  python tools/r2_profiles.py
  r2_{step,step1,step1_mask,movegen}_final_summary.csv, r2_step{,1}_final_hotlines.csv, r2_step_final_opcodes.csv,
  r2_step_final_functions.txt, r2_step_final_sass_excerpt.txt, pipes.json, traffic.json, r2_launches_final{,_by_kernel}.csv"""
import collections, csv, io, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
SO = "gym_chess_b200/libgymchess_b200.so"
PY = sys.executable


def run(*a, **env):
    e = dict(os.environ); e.update(env)
    return subprocess.run(list(a), capture_output=True, text=True, env=e)


for k in ("step_final", "movegen_final", "step1_final", "step1_mask_final"):
    r = run(PY, "tools/ncu_summary.py", "gpurun_out/r2_%s.ncu-rep" % k, "profiles/r2_%s_summary.csv" % k)
    print(k, "summary rc", r.returncode)
for rep, kern, out in (("r2_step_final", "k_env_step<2, 1, 1>", "r2_step_final_hotlines.csv"), ("r2_step1_final", "k_env_step<1, 0, 1>", "r2_step1_final_hotlines.csv")):
    r = run(PY, "tools/ncu_lines.py", "gpurun_out/%s.ncu-rep" % rep, SO, kern, "40", NCU_LINES_CSV="profiles/" + out)
    print(out, "rc", r.returncode, r.stderr[-200:])
r = run(PY, "tools/ncu_opcodes.py", "gpurun_out/r2_step_final.ncu-rep", "k_env_step<2, 1, 1>", "profiles/r2_step_final_opcodes.csv")
print("opcodes rc", r.returncode)
r = run(PY, "tools/ncu_funcs.py", "gpurun_out/r2_step_final.ncu-rep", SO, "k_env_step<2, 1, 1>")
open("profiles/r2_step_final_functions.txt", "w").write(
    "# k_env_step<2,1,1>, one 64-step launch over 524,288 envs (profiles/r2_step_final_summary.csv): executed warp instructions,\n"
    "# stall samples and active lanes per source FUNCTION (every inlined-at frame of an instruction counts for its function in the\n"
    "# inclusive table; the innermost frame only in the exclusive one) -- tools/ncu_funcs.py <rep> <so> \"k_env_step<2, 1, 1>\"\n" + r.stdout)
print("functions rc", r.returncode)
print(run(PY, "tools/ncu_pipes.py", "profiles/r2_step_final_summary.csv").returncode, "pipes/traffic")

# SASS excerpt: the window of 110 consecutive instructions with the most executed warp instructions
raw = run("ncu", "-i", "gpurun_out/r2_step_final.ncu-rep", "--page", "source", "--csv", "--print-source", "sass").stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; blocks.append(cur)
    elif row and row[0] == "Address":
        cur["hdr"] = row
    elif row and cur is not None and "hdr" in cur:
        cur["rows"].append(row)
blk = [b for b in blocks if "k_env_step<2,1,1>" in b["name"].replace("(int)", "").replace("(bool)", "").replace(" ", "")][0]
h = {n: i for i, n in enumerate(blk["hdr"])}
ex = [int(r[h["Instructions Executed"]]) for r in blk["rows"]]
th = [int(r[h["Thread Instructions Executed"]]) for r in blk["rows"]]
W = 110
best = max(range(len(ex) - W), key=lambda i: sum(ex[i:i + W]))
with open("profiles/r2_step_final_sass_excerpt.txt", "w") as f:
    f.write("# k_env_step<MODE_SAMPLED, TILE 1, SELFPLAY> (sm_100a), SASS of the %d consecutive instructions with the most executed warp\n"
            "# instructions in profiles/r2_step_final (ncu --set full, one 64-step launch over 524,288 envs; %.1f %% of the kernel's executed\n"
            "# instructions).  Pure integer code (LOP3 / IADD3 / SHF / POPC / FLO / BREV on 64-bit halves, LDS of the line masks, STS of the\n"
            "# target sets): no tensor-core or TMA instruction is expected anywhere in this path (profiles/r2_step_final_opcodes.csv has\n"
            "# the whole kernel by opcode).  columns: executed warp instructions | active lanes | SASS\n" % (W, 100.0 * sum(ex[best:best + W]) / sum(ex)))
    for i in range(best, best + W):
        f.write("%10d  %4.1f  %s\n" % (ex[i], th[i] / max(1, ex[i]), blk["rows"][i][h["Source"]]))
print("sass excerpt: instructions %d..%d" % (best, best + W))

# launch list of the bench command
src = "gpurun_out/r2_launches_final.csv"
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hd = {n: i for i, n in enumerate(rows[0])}
by = collections.OrderedDict()
with open("profiles/r2_launches_final.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel", "grid", "block", "gpu__time_duration_us"])
    for r in rows[1:]:
        name = r[hd["Kernel Name"]]
        us = float(r[hd["Metric Value"]]) / 1e3
        w.writerow([r[hd["ID"]], name[:90], r[hd["Grid Size"]], r[hd["Block Size"]], "%.2f" % us])
        key = re.sub(r"\(.*", "", name)
        by.setdefault(key, [0, 0.0]); by[key][0] += 1; by[key][1] += us
with open("profiles/r2_launches_final_by_kernel.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us"])
    for k, (n, us) in sorted(by.items(), key=lambda x: -x[1][1]):
        w.writerow([k, n, "%.1f" % us])
print("launch list: %d launches, %d kernels" % (len(rows) - 1, len(by)))

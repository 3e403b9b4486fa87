#!/usr/bin/env python
"""Attribute the SASS of one kernel to source lines (needs -lineinfo): tools/sass_lines.py <so> <kernel substring> [top]"""
import collections, os, re, subprocess, sys, tempfile

so, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
lines = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
starts = [i for i, l in enumerate(lines) if l.startswith("//--------------------- .text.")]
for si, st in enumerate(starts):
    if pat not in lines[st]:
        continue
    en = starts[si + 1] if si + 1 < len(starts) else len(lines)
    cur, cnt, tot = None, collections.Counter(), 0
    for l in lines[st:en]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
            cnt[cur] += 1
            tot += 1
    print(lines[st].strip("/- "), "instructions:", tot)
    for (f, ln), c in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
        print("  %-28s %5d %5d" % (f, ln, c))

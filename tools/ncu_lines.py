"""Joins an ncu SASS source-page export with nvdisasm line info: per CUDA source line, the share of executed warp
instructions and the average number of active threads.  Debug tooling.
  ncu -i rep --page source --csv --kernel-name regex:K --launch-skip i --launch-count 1 > sass.csv
  python tools/ncu_lines.py sass.csv cubin kernel_substring [top]"""
import csv, re, subprocess, sys
from collections import defaultdict

sass_csv, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
addr2line, cur, infn, inl = {}, None, False, ""
for l in dis:
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
_lines = open(sass_csv).read().splitlines()
import os
_starts = [i for i, l in enumerate(_lines) if l.startswith('"Address"')] + [len(_lines) + 1]
_sec = int(os.environ.get("NCU_SECTION", "0"))  # an export can hold several launches: NCU_SECTION picks one
_lines = _lines[_starts[_sec]:_starts[_sec + 1] - 1]  # (the kernel-name line may be cut mid-quote by a column filter)
rows = list(csv.reader(_lines))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iA, iE, iT, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = defaultdict(lambda: [0, 0, 0])
tot = 0
for r in rows[h + 1:]:
    if len(r) <= max(iA, iE, iT, iS) or r[0] == "Address":
        continue
    a = int(r[iA], 16)
    if base is None:
        base = a
    e, t, s = int(r[iE] or 0), int(r[iT] or 0), int(r[iS] or 0)
    line = addr2line.get(a - base, (None, ""))[0]
    agg[line][0] += e; agg[line][1] += t; agg[line][2] += s
    tot += e
print(f"total warp instructions {tot}")
srcs = {}
for (k, v) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    txt = ""
    if k:
        import glob
        f = glob.glob(f"/root/repo/**/{k[0]}", recursive=True)
        if f:
            srcs.setdefault(f[0], open(f[0]).read().splitlines())
            txt = srcs[f[0]][k[1] - 1].strip()[:100] if k[1] - 1 < len(srcs[f[0]]) else ""
    print(f"{100 * v[0] / tot:5.1f}%  thr {v[1] / max(v[0], 1):5.1f}  samples {v[2]:6d}  {k}  {txt}")

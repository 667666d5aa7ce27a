"""Dynamic opcode / pipe mix of one kernel launch from an ncu SASS source-page export (debug tooling).
  ncu -i rep --page source --csv --launch-skip i --launch-count 1 > sass.csv
  python tools/ncu_opmix.py sass.csv [top]
Pipes after /opt/skills/guides/B300_MICROARCH.md: FFMA/FMUL/FADD/IMAD/HFMA2 on the fma pipe, IADD3/LOP3/SHF/PRMT/FMNMX/... on the alu pipe."""
import csv, re, sys
from collections import defaultdict

f = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 28
_lines = open(f).read().splitlines()
import os
_starts = [i for i, l in enumerate(_lines) if l.startswith('"Address"')] + [len(_lines) + 1]
_sec = int(os.environ.get("NCU_SECTION", "0"))  # an export can hold several launches: NCU_SECTION picks one
_lines = _lines[_starts[_sec]:_starts[_sec + 1] - 1]
rows = list(csv.reader(_lines))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iE, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = defaultdict(lambda: [0, 0])
tot = tt = 0
for r in rows[h + 1:]:
    if len(r) <= max(iS, iE, iT) or r[0] == "Address":
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS].strip())
    op = m.group(2) if m else r[iS].strip()[:10]
    e, t = int(r[iE] or 0), int(r[iT] or 0)
    agg[op][0] += e; agg[op][1] += t; tot += e; tt += t
ALU = {"PRMT", "LOP3", "SHF", "ISETP", "FSETP", "FMNMX", "FMNMX3", "SEL", "IADD3", "MOV", "FSEL", "VIADD", "PLOP3", "LEA", "IABS", "POPC", "FLO", "BREV", "FCHK", "R2P", "P2R", "SGXT", "BMSK", "I2FP", "F2FP", "VIMNMX", "VIMNMX3", "CS2R"}
FMA = {"FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2"}
a = sum(v[0] for k, v in agg.items() if k in ALU)
fm = sum(v[0] for k, v in agg.items() if k in FMA)
print(f"{f}: total warp instructions {tot}, average active lanes {tt / max(tot, 1):.2f}")
print(f"  alu pipe {100 * a / tot:.1f}%  fma pipe {100 * fm / tot:.1f}%  other {100 * (tot - a - fm) / tot:.1f}%")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {k:8s} {100 * v[0] / tot:5.1f}%  lanes {v[1] / max(v[0], 1):4.1f}")

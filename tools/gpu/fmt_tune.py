"""pretty-print tools/tune_trace.py run output: python tools/gpu/fmt_tune.py < log"""
import json, sys
for l in sys.stdin:
    l = l.strip()
    p = l.split(" ", 1)
    if len(p) == 2 and p[1].startswith("{"):
        j = json.loads(p[1])
        st = j["stages"]
        print(f"{p[0]:14s} {j['scene']:8s} {j['ms_per_frame']:.4f} ms {j['mrays_s']:7.0f} Mrays/s sha {j['sha']} ext {st['ms_extend']:.3f} sh {st['ms_shade']:.3f} con {st['ms_connect']:.3f} unpiped {st['ms_frame']:.3f}")
    elif l:
        print(l[:300])

"""Per-kernel share of the step from an ncu launch list (`--metrics gpu__time_duration.sum --csv`, tools/gpu/profile.sh):
    python tools/launch_share.py profiles/r01c_launches_bench_steps2.csv [first_frame_kernel]
Per-launch times under ncu are serialised and cold-cache: the SHARES are what is comparable with bench.py."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ki, vi = rows[h].index("Kernel Name"), rows[h].index("Metric Value")
seq = [(r[ki].split("(")[0].replace("hl::", "").replace("void ", ""), float(r[vi].replace(",", "")) / 1e3) for r in rows[h + 2:] if len(r) > vi]
gen = [i for i, (k, _) in enumerate(seq) if k == "k_generate"]
frames = seq[gen[0]:]  # everything before the first k_generate is scene upload / BVH build
tot, cnt = collections.Counter(), collections.Counter()
for k, v in frames:
    tot[k] += v
    cnt[k] += 1
total = sum(tot.values())
print(f"| kernel | launches | us (sum over {len(gen)} frames) | share of the frames |\n|---|---|---|---|")
for k, v in tot.most_common():
    print(f"| {k[:60]} | {cnt[k]} | {v:.1f} | {100 * v / total:.1f} % |")
print(f"\nframe total (serialised under ncu): {total / len(gen):.1f} us per frame; build/upload kernels before the first frame: {sum(v for _, v in seq[:gen[0]]):.1f} us")

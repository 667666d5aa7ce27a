"""Markdown table of every launch in an `ncu --page raw --csv` export (one row per launch, the metrics the roofline
discussion uses).  python tools/ncu_table.py raw.csv out.md "title / command" """
import csv
import sys

METRICS = [
    ("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "warp instr"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/instr"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"), ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
]
src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
lines = [f"# {title}", "", "| # | kernel | " + " | ".join(n for m, n in METRICS if m in col) + " |", "|---|---|" + "---|" * sum(1 for m, _ in METRICS if m in col)]
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0].replace("hl::", "").replace("void ", "")
    cells = []
    for m, n in METRICS:
        if m not in col:
            continue
        v, u = r[col[m]], units[col[m]]
        try:
            f = float(v)
            if m == "gpu__time_duration.sum":
                f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                v = f"{f:.1f}"
            elif m.startswith("dram__bytes"):
                f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                v = f"{f:.1f} MB"
            elif f >= 1e6:
                v = f"{f / 1e6:.1f} M"
            else:
                v = f"{f:.4g}"
        except ValueError:
            pass
        cells.append(v)
    lines.append(f"| {k} | {name} | " + " | ".join(cells) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))

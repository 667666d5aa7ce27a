"""Summarise an `ncu --page raw --csv` export (tools/gpu/profile.sh) into profiles/: a markdown table of the metrics
the roofline discussion uses and the per-launch DRAM traffic JSON that bench.py reports as roofline.traffic.
    python tools/ncu_summary.py gpurun_out/r01b_extend_raw.csv profiles/r01b_extend k_extend "note" """
import csv
import json
import sys
from pathlib import Path

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def main():
    src, out, kernel = Path(sys.argv[1]), sys.argv[2], sys.argv[3]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if kernel in r[rows[0].index("Kernel Name")]]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {Path(out).name} — ncu `--set full` capture of {kernel} ({len(data)} launches)", "", note, "", "| metric | " + " | ".join(f"launch {k}" for k in range(len(data))) + " |", "|---|" + "---|" * len(data)]
    for m in METRICS:
        if m in col:
            lines.append(f"| {m} | " + " | ".join(f"{r[col[m]]} {units[col[m]]}".strip() for r in data) + " |")
    Path(out + "_ncu_full.md").write_text("\n".join(lines) + "\n")

    def mb(r, m):
        v, u = float(r[col[m]]), units[col[m]]
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]

    per = [mb(r, "dram__bytes_read.sum") + mb(r, "dram__bytes_write.sum") for r in data]
    us = [float(r[col["gpu__time_duration.sum"]]) for r in data]
    # mean over the launches that carry the step (a capture of an almost empty queue would only dilute it)
    big = [p for p, t in zip(per, us) if t > 0.05 * max(us)]
    def pct(m):
        return [float(r[col[m]]) for r in data] if m in col else None

    json.dump({"kernel": kernel, "launches": len(per), "dram_bytes_per_launch_mb": per, "duration_us": us, "mean_mb": sum(big) / len(big), "mean_over": len(big), "source": out + "_ncu_full.md",
               # what actually bounds the kernel (first captured launch = the largest one)
               "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"), "lanes_per_instruction": pct("smsp__thread_inst_executed_per_inst_executed.ratio"),
               "alu_pipe_pct": pct("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), "fma_pipe_pct": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
               "dram_throughput_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "l2_throughput_pct": pct("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "l1_hit_pct": pct("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": pct("lts__t_sector_hit_rate.pct"), "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active")},
              open(out + "_traffic.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()

"""Summarise an ncu report (.ncu-rep, read on the CPU box) and a launch list (csv) into profiles/<name>.md.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv profiles/r01_k1.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main(rep, launches, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = ["# ncu summary of `%s`" % rep, "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (B200); read here with",
             "`ncu -i ... --page raw --csv`.  One column per captured launch.", ""]
    lines.append("| metric | unit | " + " | ".join(r[ki].split("(")[0][-40:] for r in data) + " |")
    lines.append("|---|---|" + "---|" * len(data))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append("| `%s` | %s | " % (k, units[i]) + " | ".join(r[i] for r in data) + " |")
    if launches:
        d = defaultdict(list)
        rr = [r for r in csv.reader(open(launches)) if len(r) > 5]
        h = rr[0]
        kn, mv = h.index("Kernel Name"), h.index("Metric Value")
        for r in rr[1:]:
            try:
                d[r[kn].split("(")[0]].append(float(r[mv].replace(",", "")))
            except ValueError:
                pass
        lines += ["", "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised)", "",
                  "| kernel | launches | avg us | min us | max us |", "|---|---|---|---|---|"]
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            lines.append("| `%s` | %d | %.1f | %.1f | %.1f |" % (k, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None, sys.argv[3])

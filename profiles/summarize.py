"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) and raw pages into markdown.

    python profiles/summarize.py launches <launches.csv>
    python profiles/summarize.py raw <raw_page.csv> [...]
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        agg.setdefault(row["Kernel Name"].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | total ms | share |\n|---|---:|---:|---:|---:|")
    for k, v in agg.items():
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e6:.2f} | {sum(v) / tot:.3f} |")


def raw(paths):
    for p in paths:
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:3]:
            print(f"\n**{r[hdr.index('Kernel Name')]}** ({p})\n\n| metric | value | unit |\n|---|---:|---|")
            for k in KEYS:
                if k in hdr:
                    print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2:])

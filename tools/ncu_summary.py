#!/usr/bin/env python
"""Condense the CSV pages tools/ncu_capture.sh exports (raw + source) into the handful of numbers the design notes
quote: duration, registers, occupancy limiters, DRAM traffic / throughput, issue utilisation, pipe utilisation, stall
reasons per issue, opcode histogram and executed warp-instructions.

    python tools/ncu_summary.py gpurun_out/<name> [agent_steps_per_launch]
"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main():
    base = sys.argv[1]
    rows = list(csv.reader(open(base + "_raw.csv")))
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))
    print("kernel:", d.get("Kernel Name"))
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {u.get(k, '')}")
    for k in sorted(d):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            v = float(d[k] or 0)
            if v >= 0.15:
                print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} = {v:.2f}")
    try:
        rows = list(csv.reader(open(base + "_source.csv")))
    except OSError:
        return
    hdr, data = rows[1], rows[2:]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ops, stl = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
        if m:
            ops[m.group(2).split(".")[0]] += int(r[ie] or 0)
            stl[m.group(2).split(".")[0]] += int(r[isamp] or 0)
    tot, ts = sum(ops.values()), max(1, sum(stl.values()))
    print(f"  executed warp-instructions = {tot}  (static SASS instructions {len(data)})")
    if len(sys.argv) > 2:
        print(f"  thread-instructions per agent-step = {32.0 * tot / float(sys.argv[2]):.1f}")
    print("  opcodes: " + ", ".join(f"{o} {100.0 * c / tot:.1f}% (stall {100.0 * stl[o] / ts:.0f}%)" for o, c in ops.most_common(14)))


if __name__ == "__main__":
    main()

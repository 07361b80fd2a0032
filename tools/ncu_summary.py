#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_composite.ncu-rep profiles/r1_ncu_composite.txt
"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg",
]
STALLS = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_drain", "stall_lg", "stall_long_sb",
          "stall_math", "stall_membar", "stall_mio", "stall_misc", "stall_no_inst", "stall_not_selected",
          "stall_selected", "stall_short_sb", "stall_sleep", "stall_tex", "stall_wait"]


def ncu(rep, *args):
    r = subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True)
    return r.stdout


def to_int(x):
    try:
        return int(x)
    except Exception:
        return 0


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = [f"# summary of {rep} (ncu --set full --clock-control none --import-source on; per launch)"]
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for r in rows[2:]:
        lines.append("")
        lines.append(f"## {r[ik]}")
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"{m:90s} {r[i]:>16s} {units[i]}")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    cur, data, h = None, {}, None
    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            cur = r[1]
            data.setdefault(cur, [])
        elif r[0] == "Line No":
            h = r
        elif cur and r[0].isdigit():
            data[cur].append(r)
    if h:
        iI, iS, iW, iE = (h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared"),
                          h.index("L1 Wavefronts Shared Excessive"))
        for k, v in data.items():
            ti = sum(to_int(r[iI]) for r in v) or 1
            ts = sum(to_int(r[iS]) for r in v) or 1
            tw = sum(to_int(r[iW]) for r in v) or 1
            lines.append("")
            lines.append(f"## source hot spots: {k}")
            ss = {s: sum(to_int(r[h.index(s)]) for r in v) for s in STALLS if s in h}
            t = sum(ss.values()) or 1
            lines.append("warp-state samples: " + ", ".join(f"{s[6:]} {x / t * 100:.1f}%" for s, x in
                                                             sorted(ss.items(), key=lambda kv: -kv[1]) if x / t > 0.01))
            lines.append(f"{'line':>5} {'inst%':>6} {'samp%':>6} {'smem wf%':>8} {'excess%':>7}  source")
            for r in v:
                a, b, c = to_int(r[iI]) / ti, to_int(r[iS]) / ts, to_int(r[iW]) / tw
                if a > 0.01 or b > 0.015 or c > 0.02:
                    lines.append(f"{r[0]:>5} {a * 100:6.2f} {b * 100:6.2f} {c * 100:8.2f} {to_int(r[iE]) / tw * 100:7.2f}  "
                                 f"{r[1].strip()[:100]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"wrote {out} ({len(lines)} lines)")


if __name__ == "__main__":
    main()

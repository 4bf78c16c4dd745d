"""Pipe utilisation and warp-stall breakdown of a kernel from an .ncu-rep (reads it with ncu --page raw --csv).
usage: python tools/ncu_stalls.py <report.ncu-rep>"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for v in rows[2:]:
    d = dict(zip(h, v))
    print(d.get("Kernel Name"))
    for k in ("gpu__time_duration.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "launch__registers_per_thread",
              "dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k in d: print("  %-75s %s" % (k, d[k]))
    st = sorted(((float(b), a) for a, b in d.items() if "issue_stalled" in a and a.endswith("per_issue_active.ratio") and b not in ("", "n/a")), reverse=True)
    for b, a in st[:8]:
        print("  %-75s %.3f" % (a.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), b))

"""Summarise gpurun_out ncu output into profiles/ (launch-list shares, key metrics of the full captures)."""
import csv, collections, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = os.path.join(ROOT, "gpurun_out", sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r2p")
P = os.path.join(ROOT, "profiles")

# ---- launch list ----
lines = [l for l in open(os.path.join(G, "launches.csv")) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[row["Metric Unit"]]
    k = re.sub(r"\(.*", "", row["Kernel Name"])
    k = re.sub(r"void |swg::|\(anonymous namespace\)::|<unnamed>::", "", k)
    k = re.sub(r"\((?:int|bool)\)", "", k)
    fam = re.sub(r"wavefront_kernel<(Lane\d+), (\d+), (\d+), (\d), (\d), (\d+), (\d+)>", r"wavefront_kernel<\1,G=\2,K=*,MP=\4,GP=\5,imm=\6/\7>", k)
    fam = re.sub(r"wavefront_q2_kernel<(\d+), (\d+), (\d), (\d), (\d+), (\d+)>", r"wavefront_q2_kernel<G=\1,K=*,CIN=\3,COUT=\4,imm=\5/\6>", fam)
    fam = re.sub(r"wavefront_xw_kernel<(Lane\d+), (\d+), (\d+), (\d+)>", r"wavefront_xw_kernel<\1,K=*,imm=\3/\4>", fam)
    tot[fam] += v; cnt[fam] += 1
T = sum(tot.values())
with open(os.path.join(P, tag + "_launch_list_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu-baseline --no-pipebench --no-extra\n")
    f.write("(per-launch times are cold-cache and serialised: compare shares)\n\n%-64s %6s %12s %7s\n" % ("kernel", "n", "ms", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        f.write("%-64s %6d %12.3f %6.2f%%\n" % (k[:64], cnt[k], v / 1e6, 100 * v / T))
print(open(os.path.join(P, tag + "_launch_list_summary.txt")).read())

# ---- full captures ----
keys = ["gpu__time_duration.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
res = {}
keys += ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for i, name in [(1, "q2_first_pass"), (2, "q2_middle_pass"), (3, "q2_last_pass"), (4, "seqpair_q144"), (5, "xw_cfg4_q5478")]:
    rep = os.path.join(G, "prof3_%d.ncu-rep" % i)
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, d = rows[0], rows[1], rows[2]
    idx = {h: j for j, h in enumerate(hdr)}
    res[name] = {"kernel": d[idx["Kernel Name"]], **{k: (d[idx[k]] + " " + units[idx[k]]).strip() for k in keys if k in idx}}
json.dump(res, open(os.path.join(P, tag + "_ncu_full_cfg2_kernels.json"), "w"), indent=1)
for k, v in res.items():
    print(k, v["kernel"], v["gpu__time_duration.sum"], "ALU", v["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"],
          "issue", v["smsp__issue_active.avg.pct_of_peak_sustained_active"], "dram rd", v["dram__bytes_read.sum"], "wr", v["dram__bytes_write.sum"])

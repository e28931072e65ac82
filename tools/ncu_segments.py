"""Development probe: per-region instruction / stall shares of one kernel from an ncu report (source page, SASS).
    python tools/ncu_segments.py gpurun_out/x.ncu-rep <kernel regex> <units per launch>"""
import csv, subprocess, sys
rep, kern, N = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; r = rows[2]
for m in ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"] + \
         [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]:
    if m in hdr:
        v = r[hdr.index(m)]
        try:
            if float(v) == 0: continue
        except ValueError: pass
        print(f"{m:95s} {v} {rows[1][hdr.index(m)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
ie = hdr.index("Instructions Executed"); isrc = hdr.index("Source"); iss = hdr.index("Warp Stall Sampling (All Samples)")
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ie].isdigit()]
tot = sum(int(r[ie]) for r in data); ts = sum(int(r[iss]) for r in data)
print("SASS instrs", len(data), "executed/unit", tot / N)
lvl = lambda i: int(data[i][ie])
s = 0
for i in range(1, len(data) + 1):
    if i == len(data) or abs(lvl(i) - lvl(i - 1)) > 0.3 * max(lvl(i), lvl(i - 1), 1):
        ex = sum(lvl(k) for k in range(s, i)); st = sum(int(data[k][iss]) for k in range(s, i))
        if ex > 0.01 * tot or st > 0.01 * ts:
            print(f"[{s:5d},{i:5d}) n={i-s:4d} exec/unit={ex/N:7.1f} ({100*ex/tot:4.1f}%) stall {100*st/ts:4.1f}%  lvl/unit={lvl(s)/N:6.2f}  first: {data[s][isrc][:60]}")
        s = i

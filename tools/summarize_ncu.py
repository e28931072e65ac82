"""Summarises gpurun_out/<tag>_full.ncu-rep (+ <tag>_launches.csv) into profiles/<tag>_ncu_summary.md and
profiles/r1_traffic.json (DRAM bytes per launch of the hot kernels, read by bench.py).
    python tools/summarize_ncu.py r1g "commit / description"
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rep = ROOT / "gpurun_out" / f"{tag}_full.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
short = lambda n: n.split("(")[0].replace("void locohd::<unnamed>::", "").replace("locohd::<unnamed>::", "")
metrics = [
    ("gpu__time_duration.sum", "duration (ms)"), ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data-pipe wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]
units = rows[1]
ki = hdr.index("Kernel Name")
kernels = rows[2:]
out = [f"# {tag} — ncu `--set full --clock-control none` on B200 ({note})", "",
       "Command: `ncu --set full --clock-control none --import-source on -k regex:\"env_fused|score_fast|env_tile|"
       "build_cells\" -s 8 -c 4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline`",
       "(default workload: 256 structure pairs of config-2 shape = 5.12 M environments, 2.56 M anchor pairs per launch)", ""]
names = [short(r[ki]) for r in kernels]
out.append("| metric | " + " | ".join(names) + " |")
out.append("|---|" + "---|" * len(names))
traffic = {}
group = {"env_tile_kernel<0>": "count", "env_tile_kernel<1>": "fill", "env_fused_kernel": "fill",
         "env_sort_kernel<256, 0>": "sort", "score_fast_kernel": "score", "build_cells_kernel": "cells"}
for m, label in metrics:
    if m not in hdr:
        continue
    i = hdr.index(m)
    out.append(f"| {label} ({units[i]}) | " + " | ".join(r[i][:12] for r in kernels) + " |")
ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
for r, n in zip(kernels, names):
    b = float(r[ri]) * scale.get(units[ri], 1.0) + float(r[wi]) * scale.get(units[wi], 1.0)
    for k, g in group.items():
        if n.startswith(k) or k in n:
            traffic.setdefault(g, b)   # one launch per group (the capture holds one step)
            break
lf = ROOT / "gpurun_out" / f"{tag}_launches.csv"
if lf.exists():
    lr = [r for r in csv.reader(open(lf)) if len(r) > 10]
    h = lr[0]
    k2, v2 = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in lr[1:]:
        try:
            v = float(r[v2].replace(",", ""))
        except ValueError:
            continue
        agg[short(r[k2])][0] += 1
        agg[short(r[k2])][1] += v
    tot = sum(v[1] for k, v in agg.items() if "fp64_peak" not in k)
    out += ["", "Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 ... bench.py --steps 2 "
            "--warmup 3 --no-cpu-baseline`: warm-up, timed, work-count and e2e passes; cold-cache serialised times — "
            "compare shares):", "",
            "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if "fp64_peak" in n:
            continue
        out.append(f"| {n} | {c} | {v / 1e6:.3f} | {100 * v / tot:.1f}% |")
(ROOT / "profiles" / f"{tag}_ncu_summary.md").write_text("\n".join(out) + "\n")
# issue-slot utilisation and instruction counts of the same launches (bench.py quotes them beside the rooflines)
pipes = {}
for r, n in zip(kernels, names):
    for k, g in group.items():
        if (n.startswith(k) or k in n) and g not in pipes:
            def val(m):
                return float(r[hdr.index(m)]) if m in hdr else None
            pipes[g] = {"warp_instructions": val("smsp__inst_executed.sum"),
                        "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "shared_pipe_busy_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                        "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")}
            break
(ROOT / "profiles" / "r1_traffic.json").write_text(json.dumps(
    {"workload": "cfg2", "anchor_pairs_per_step": 2560000, "source": f"profiles/{tag}_ncu_summary.md",
     "dram_bytes_per_launch": traffic, "pipes": pipes}, indent=1) + "\n")
print("\n".join(out[:40]))
print(traffic)

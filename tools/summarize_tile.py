"""Summarises the ncu captures of the tile scoring kernel (tools/r6_evidence.sh) under gpurun_out/ into profiles/:
    <tag>_full.ncu-rep        ncu --set full on score_tile_kernel, `bench.py --workload cfg5 --ensemble 48 --steps 1 --warmup 3`
    <tag>_cfg5_metrics.csv    selected metrics of the kernels of the DEFAULT workload (1000-structure ensemble)
    <tag>_launches.csv        every launch of `bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline` with its time
-> profiles/<tag>_ncu_summary.md, profiles/<tag>_cfg5_traffic.json (read by bench.py).
    python tools/summarize_tile.py r6f "note"
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
G = ROOT / "gpurun_out"
short = lambda n: n.split("(")[0].replace("void locohd::<unnamed>::", "").replace("locohd::<unnamed>::", "").replace("void unnamed>::", "").replace("unnamed>::", "")
GROUP = {"env_tile_kernel<0>": "count", "env_fused_kernel": "fill", "score_tile_kernel": "score", "score_fast_kernel": "score", "build_cells_kernel": "cells"}
METRICS = [
    ("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers / thread"),
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
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
out = [f"# {tag} — ncu on B200 ({note})", ""]


def group_of(name):
    for k, g in GROUP.items():
        if k in name:
            return g
    return None


# ---- cfg2: full set
raw = subprocess.run(["ncu", "-i", str(G / f"{tag}_full.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, kernels = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
names = [short(r[ki]) for r in kernels]
out += ["## score_tile_kernel, full set (all-vs-all ensemble of 48 structures: 1 128 structure pairs, 5.64 M anchor pairs per launch)", "",
        "`ncu --set full --clock-control none --import-source on -k regex:score_tile -s 3 -c 1 "
        "python bench.py --workload cfg5 --ensemble 48 --steps 1 --warmup 3 --no-cpu-baseline --no-extras`", "",
        "| metric | " + " | ".join(names) + " |", "|---|" + "---|" * len(names)]
traffic2, pipes2 = {}, {}
for m, label in METRICS:
    if m in hdr:
        i = hdr.index(m)
        out.append(f"| {label} ({units[i]}) | " + " | ".join(r[i][:12] for r in kernels) + " |")
ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
for r, n in zip(kernels, names):
    g = group_of(n)
    if g and g not in traffic2:
        traffic2[g] = float(r[ri]) * SCALE.get(units[ri], 1.0) + float(r[wi]) * SCALE.get(units[wi], 1.0)
        pipes2[g] = {"warp_instructions": float(r[hdr.index("smsp__inst_executed.sum")]),
                     "issue_slots_busy_pct": float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
                     "shared_pipe_busy_pct": float(r[hdr.index("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed")]),
                     "fp64_pipe_pct": float(r[hdr.index("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")])}
for r, n in zip(kernels, names):
    if group_of(n) == "score":
        P48 = 5.64e6
        g = lambda m: float(r[hdr.index(m)])
        out += ["", f"Per anchor pair: {g('smsp__inst_executed.sum') / P48:.0f} warp instructions, "
                f"{g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / P48:.0f} shared-memory wavefronts "
                f"({g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / P48:.0f} of them bank conflicts); "
                "one pair per warp (score_fast_kernel, profiles/r5c): 1 949 instructions, 523 wavefronts, 11.61 ms for the same launch."]
        stalls = sorted(((float(r[i]), n2) for i, n2 in enumerate(hdr) if "issue_stalled" in n2 and "per_issue_active" in n2 and "not_issued" not in n2
                         and r[i] not in ("", "n/a")), reverse=True)[:7]
        out += ["", "Warps stalled per issued instruction: " + ", ".join(f"{n2.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, n2 in stalls)]
# ---- cfg5 (default workload): selected metrics
rows = [r for r in csv.reader(open(G / f"{tag}_cfg5_metrics.csv")) if len(r) > 10]
h = rows[0]
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault((r[h.index("ID")], short(r[h.index("Kernel Name")])), {})[r[h.index("Metric Name")]] = (float(r[h.index("Metric Value")].replace(",", "")), r[h.index("Metric Unit")])
out += ["", "## Default workload: all-vs-all ensemble of 1000 structures (2.4975e9 anchor pairs, 5.0 M environments per launch)", "",
        "`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,… --clock-control none "
        "-k regex:\"score_tile|env_fused|build_cells|env_tile\" -s 12 -c 4 python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline`", ""]
mnames = list(next(iter(per.values())).keys())
out += ["| metric | " + " | ".join(k[1] for k in per) + " |", "|---|" + "---|" * len(per)]
for m in mnames:
    out.append(f"| {m} ({next(iter(per.values()))[m][1]}) | " + " | ".join(f"{v[m][0]:.6g}" for v in per.values()) + " |")
traffic5, pipes5 = {}, {}
for (_, n), v in per.items():
    g = group_of(n)
    if g and g not in traffic5:
        traffic5[g] = v["dram__bytes_read.sum"][0] + v["dram__bytes_write.sum"][0]
        pipes5[g] = {"warp_instructions": v["smsp__inst_executed.sum"][0],
                     "issue_slots_busy_pct": v["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
                     "shared_pipe_busy_pct": v["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][0],
                     "fp64_pipe_pct": v["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][0],
                     "shared_wavefronts": v["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0],
                     "shared_bank_conflicts": v["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"][0]}
        if g == "score":
            pairs = 2497500000.0
            out += ["", f"K2 per anchor pair: {v['smsp__inst_executed.sum'][0] / pairs:.0f} warp instructions, "
                    f"{v['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'][0] / pairs:.0f} shared-memory wavefronts "
                    f"({v['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'][0] / pairs:.0f} of them bank conflicts), "
                    f"{v['dram__bytes_read.sum'][0] / pairs:.0f} B read from DRAM (8 environments staged per 16 pairs = ~760 B per pair), "
                    f"{v['dram__bytes_write.sum'][0] / pairs:.1f} B written."]
(ROOT / "profiles" / f"{tag}_cfg5_traffic.json").write_text(json.dumps(
    {"workload": "cfg5", "anchor_pairs_per_step": 2497500000, "source": f"profiles/{tag}_ncu_summary.md",
     "dram_bytes_per_launch": traffic5, "pipes": pipes5}, indent=1))

# ---- launch list of the default command
lf = G / f"{tag}_launches.csv"
if lf.exists():
    rows = [r for r in csv.reader(open(lf)) if len(r) > 10]
    h = rows[0]
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = short(r[h.index("Kernel Name")])
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[h.index("Metric Value")].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    out += ["", "## Launch list of the default command", "",
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 … python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline` "
            "(warm-up, timed, work-count and e2e passes; cold-cache serialised times — compare shares):", "",
            "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {n} | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.2f}% |")
    (ROOT / "profiles" / f"{tag}_launches.csv").write_text(lf.read_text())
(ROOT / "profiles" / f"{tag}_ncu_summary.md").write_text("\n".join(out) + "\n")
print("\n".join(out))

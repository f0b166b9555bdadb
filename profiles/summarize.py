"""Turns gpurun_out/launches_<tag>.csv and gpurun_out/step_<tag>.ncu-rep into the tracked summaries
profiles/<tag>_launches.md and profiles/<tag>_kernels.md.   Usage: python profiles/summarize.py <tag>"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

tag = sys.argv[1]
root = Path(__file__).resolve().parent.parent
out = root / "profiles"

rows = list(csv.reader(open(root / "gpurun_out" / f"launches_{tag}.csv")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg.setdefault(r[ki].split("(")[0].replace("void ", "").replace("tq::", ""), []).append(v)
tot = sum(sum(v) for v in agg.values())
lines = [f"# ncu launch list `{tag}` — `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (C2, 100k units/step)",
         "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)", "",
         "| kernel | launches | avg us | share |", "|---|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f} % |")
(out / f"{tag}_launches.md").write_text("\n".join(lines) + "\n")

raw = subprocess.run(["ncu", "-i", str(root / "gpurun_out" / f"step_{tag}.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
lines = [f"# ncu --set full `{tag}` (one launch per kernel, C2 workload: 100 000 units)", ""]
cols = [i for i, c in enumerate(hh) if c in want]
names = [r[hh.index("Kernel Name")].split("(")[0].replace("void ", "").replace("tq::", "") for r in rr[2:]]
lines.append("| metric | unit | " + " | ".join(f"`{n[:40]}`" for n in names) + " |")
lines.append("|---|---|" + "---:|" * len(names))
for i in cols:
    if hh[i] == "Kernel Name":
        continue
    lines.append(f"| {hh[i]} | {rr[1][i]} | " + " | ".join(r[i] for r in rr[2:]) + " |")
(out / f"{tag}_kernels.md").write_text("\n".join(lines) + "\n")
print((out / f"{tag}_launches.md").read_text())
print((out / f"{tag}_kernels.md").read_text())

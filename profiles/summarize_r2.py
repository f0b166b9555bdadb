"""Turns gpurun_out/launches_r2_<w>.csv and gpurun_out/*_raw.csv (ncu --page raw --csv exports made on the GPU box by
profiles/capture_r2.sh) into the tracked summaries profiles/r2_launches_<w>.md / profiles/r2_kernels_<name>.md."""
import collections
import csv
import json
import sys
from pathlib import Path

root = Path(__file__).resolve().parent.parent
out = root / "profiles"
go = root / "gpurun_out"

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("tq::", "")


def launches(w, what):
    rows = list(csv.reader(open(go / f"launches_r2_{w}.csv")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(short(r[ki]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    lines = [f"# ncu launch list r2, {what}", "",
             f"`python bench.py --workload {w} --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0` under "
             "`ncu --metrics gpu__time_duration.sum --clock-control none` (three steps; cold-cache, serialised: compare SHARES)", "",
             "| kernel | launches | avg us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f} % |")
    (out / f"r2_launches_{w}.md").write_text("\n".join(lines) + "\n")
    return {k: sum(v) / len(v) for k, v in agg.items()}


def kernels(csv_name, md_name, title):
    rr = list(csv.reader(open(go / csv_name)))
    hh = rr[0]
    names = [short(r[hh.index("Kernel Name")]) for r in rr[2:]]
    lines = [f"# ncu --set full r2: {title}", "", "| metric | unit | " + " | ".join(f"`{n[:44]}`" for n in names) + " |",
             "|---|---|" + "---:|" * len(names)]
    got = {}
    for m in WANT:
        if m in hh:
            i = hh.index(m)
            lines.append(f"| {m} | {rr[1][i]} | " + " | ".join(r[i] for r in rr[2:]) + " |")
            got[m] = [r[i] for r in rr[2:]]
    (out / md_name).write_text("\n".join(lines) + "\n")
    return names, got, rr[1], hh


if __name__ == "__main__":
    launches("c3", "C3 (1000 AOIs x 5000 frames = 5 000 000 units per step)")
    launches("c2", "C2 (100 AOIs x 1000 frames = 100 000 units per step)")
    traffic = {}
    for w, units in (("c3", 5_000_000), ("c2", 100_000)):
        names, got, units_row, hh = kernels(f"step_r2_{w}_raw.csv", f"r2_kernels_{w}.md",
                                            f"one launch per kernel, workload {w} ({units} units per launch)")
        for j, n in enumerate(names):
            if n.startswith("ksmogn_stream_kernel"):
                def val(m):
                    v, u = float(got[m][j].replace(",", "")), units_row[hh.index(m)]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
                traffic[f"ksmogn_stream_kernel_{w}"] = {
                    "workload": w, "units_per_launch": units, "dram_bytes_read": val("dram__bytes_read.sum"),
                    "dram_bytes_write": val("dram__bytes_write.sum"),
                    "source": f"profiles/r2_kernels_{w}.md (ncu --set full, capture_r2.sh, single-bin form)"}
    kernels("ksmogn_o64_r2_raw.csv", "r2_kernels_o64.md", "likelihood kernel, C2 with a 64-bin offset histogram (many-bins form)")
    kernels("ksmogn_o3_r2_raw.csv", "r2_kernels_o3.md", "likelihood kernel, C2 with the simulator's three bins kept distinct (register-cached form)")
    old = json.loads((out / "traffic.json").read_text()) if (out / "traffic.json").exists() else {}
    old.update(traffic)
    (out / "traffic.json").write_text(json.dumps(old, indent=1) + "\n")
    for f in ("r2_launches_c3.md", "r2_launches_c2.md", "r2_kernels_c3.md", "r2_kernels_o64.md"):
        print((out / f).read_text())

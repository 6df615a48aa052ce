#!/usr/bin/env python
"""Turn one GPU-box visit's ncu outputs into the tracked summary under profiles/.

    python tools/summarize_profile.py TAG          (reads gpurun_out/launches_TAG.csv, prof_*_TAG.ncu-rep,
                                                    bench_TAG.json; writes profiles/TAG_*.{md,csv})
"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

TAG = sys.argv[1]
OUT = "profiles"
os.makedirs(OUT, exist_ok=True)
md = [f"# ncu summary `{TAG}`", ""]

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor inst %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall: dispatch"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: mio throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
]

launches = f"gpurun_out/launches_{TAG}.csv"
if os.path.exists(launches):
    shutil.copy(launches, f"{OUT}/{TAG}_launches.csv")
    hdr, rows = None, []
    for r in csv.reader(open(launches)):
        if r and r[0] == "ID":
            hdr = r
        elif hdr and r and r[0].isdigit():
            rows.append(r)
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0, set()])
        a[0] += 1; a[1] += float(r[vi].replace(",", "")); a[2].add(r[gi])
    tot = sum(v[1] for v in agg.values())
    md += [f"## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, {len(rows)} launches; "
           "cold-cache, serialised: read the SHARES)", "", "| kernel | launches | total us | share | grids |", "|---|---:|---:|---:|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | {' '.join(sorted(v[2]))[:60]} |")
    md.append("")

for rep in sorted(glob.glob(f"gpurun_out/prof_*_{TAG}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    base = os.path.basename(rep)[:-8]
    with open(f"{OUT}/{base}_raw_selected.csv", "w") as f:
        w = csv.writer(f)
        cols = ["Kernel Name", "Grid Size", "Block Size"] + [k for k, _ in KEYS if k in idx]
        w.writerow(cols); w.writerow([units[idx[c]] for c in cols])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])
    md += [f"## `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`)", ""]
    for r in rows[2:]:
        md.append(f"### `{r[idx['Kernel Name']][:90]}` grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        md += ["", "| metric | value |", "|---|---:|"]
        for k, label in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                md.append(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |")
        md.append("")

# measured DRAM traffic of the GEMM family for ONE step (bench.py reports it as roofline.traffic): all launches of the
# tensor-core capture of this visit, which gpu_prof.sh takes over exactly one step
tc_rep = f"gpurun_out/prof_tc_{TAG}.ncu-rep"
if os.path.exists(tc_rep):
    raw = subprocess.run(["ncu", "-i", tc_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) >= 3:
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per, total = {}, 0.0
        for n, r in enumerate(rows[2:]):
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]]
            total += rd + wr
            per[f"{n}: {r[idx['Kernel Name']][:70]} grid {r[idx['Grid Size']]}"] = {
                "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6,
                "duration_us": float(r[idx["gpu__time_duration.sum"]].replace(",", "")),
                "tensor_pipe_active_pct": float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])}
        meta = {}
        try:
            meta = json.load(open(f"gpurun_out/bench_{TAG}.json"))["config"]
        except Exception:
            pass
        gemm = "fp16x2" if "2-term fp16" in meta.get("gemm", "") else ("tf32x3" if "tf32" in meta.get("gemm", "") else
                                                                        ("bf16" if "bf16" in meta.get("gemm", "") else "simt"))
        json.dump({"source": f"profiles/prof_tc_{TAG}_raw_selected.csv (ncu --set full --clock-control none, dram__bytes_read.sum + "
                             "dram__bytes_write.sum, the tensor-core GEMM launches of one step)",
                   "workload": meta.get("workload", "humanoid-cfg3"), "gemm": gemm, "family_bytes_per_step": total,
                   "launches": len(per), "per_launch": per}, open(f"{OUT}/traffic_{gemm}.json", "w"), indent=1)
        md += [f"## DRAM traffic of the GEMM family, one step: {total / 1e9:.3f} GB over {len(per)} launches "
               f"(`profiles/traffic_{gemm}.json`)", ""]

bench = f"gpurun_out/bench_{TAG}.json"
if os.path.exists(bench) and os.path.getsize(bench):
    shutil.copy(bench, f"{OUT}/{TAG}_bench.json")
    d = json.load(open(bench))
    md += ["## bench line of the same visit (CUDA-event timing, not under the profiler)", "", "```json",
           json.dumps({k: d[k] for k in ("value", "ms_per_step", "roofline", "e2e", "gpu_launches", "clocks") if k in d}, indent=1),
           "```", ""]
open(f"{OUT}/{TAG}_summary.md", "w").write("\n".join(md))
print("wrote", f"{OUT}/{TAG}_summary.md")

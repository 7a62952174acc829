#!/usr/bin/env python3
"""Turn gpurun_out/ scratch (ncu launch list, ncu --set full report, sweeps) into the tracked
summaries under profiles/.  Usage: python tools/summarize_profiles.py r01 [workload]"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
workload = sys.argv[2] if len(sys.argv) > 2 else "1m_fp32_q1_top10"
os.makedirs(PROF, exist_ok=True)

# ---- launch list -----------------------------------------------------------------------------
for lp, workload_ in ((os.path.join(OUT, "launches.csv"), workload), (os.path.join(OUT, "launches_gemm.csv"), "10m_bf16_q256_top100")):
    if not os.path.exists(lp):
        continue
    rows = list(csv.reader(open(lp)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    d = defaultdict(list)
    order = []
    for r in rows[h + 1:]:
        if len(r) > vi:
            d[r[ki]].append(float(r[vi].replace(",", "")) / 1e3)
            order.append((r[ki], float(r[vi].replace(",", "")) / 1e3))
    with open(os.path.join(PROF, f"{tag}_launches_{workload_}.md"), "w") as f:
        f.write(f"# ncu launch list -- bench.py --steps 5 (3 for the tensor-core workload) --warmup 3 --reps 1 ({workload_}), {tag}\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: "
                "compare shares, not absolutes).\n\n| kernel | launches | mean us | min us | max us |\n|---|---:|---:|---:|---:|\n")
        for k, v in d.items():
            f.write(f"| `{k[:70]}` | {len(v)} | {sum(v)/len(v):.1f} | {min(v):.1f} | {max(v):.1f} |\n")
        step = [(k, t) for k, t in order if "cab_scan" in k or "gemv_scan" in k or "finalize" in k or "emit" in k or "gemm" in k]
        if step:
            tot = sum(t for _, t in step)
            f.write("\nShare of the search step (scan + finalize [+ emit]) by kernel:\n\n")
            agg = defaultdict(float)
            for k, t in step:
                agg[k.split("(")[0]] += t
            for k, t in agg.items():
                f.write(f"* `{k}`: {100 * t / tot:.1f} %\n")
    print("wrote launches summary")

# ---- full capture ------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
traffic = {}
REPORT_WORKLOAD = {"prof_gemv": "1m_fp32_q1_top10", "prof_gemm": "10m_bf16_q256_top100",
                   "prof_gemv10m": "10m_fp32_q1_top10", "prof_finalize": "finalize_1m_fp32_q1_top10",
                   "prof_score_all": "score_all_10m_fp32"}
for rep in sorted(f for f in os.listdir(OUT) if f.endswith(".ncu-rep")):
    r = subprocess.run(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        continue
    H, U = rows[0], rows[1]
    name = rep.replace(".ncu-rep", "")
    wl = REPORT_WORKLOAD.get(name, workload)
    with open(os.path.join(PROF, f"{tag}_{name}_{wl}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none: {name} ({wl}), {tag}\n\n")
        for r_ in rows[2:]:
            kn = r_[H.index("Kernel Name")]
            f.write(f"## `{kn}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in H:
                    f.write(f"| {w} | {r_[H.index(w)]} | {U[H.index(w)]} |\n")
            f.write("\n")
            try:
                rd = float(r_[H.index("dram__bytes_read.sum")].replace(",", ""))
                wr = float(r_[H.index("dram__bytes_write.sum")].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
                tb = rd * scale[U[H.index("dram__bytes_read.sum")]] + wr * scale[U[H.index("dram__bytes_write.sum")]]
                traffic.setdefault((wl, kn.split("(")[0]), []).append(tb)
            except Exception:
                pass
    print("wrote", name)
if traffic:
    tp = os.path.join(PROF, "traffic.json")
    t = json.load(open(tp)) if os.path.exists(tp) else {}
    for (wl, kern), vals in traffic.items():
        t[wl] = sum(vals) / len(vals)
        t[wl + "_kernel"] = kern
    json.dump(t, open(tp, "w"), indent=1)
    print("traffic", t)

# ---- sweeps ------------------------------------------------------------------------------------
for name in ("sweep_1m.jsonl", "sweep_10m.jsonl"):
    p = os.path.join(OUT, name)
    if os.path.exists(p):
        recs = [json.loads(l) for l in open(p) if l.startswith("{")]
        with open(os.path.join(PROF, f"{tag}_gemv_{name.replace('.jsonl', '')}.md"), "w") as f:
            f.write(f"# GEMV scan configuration sweep ({name}), {tag}\n\nCUDA events around the scan kernel on its "
                    "stream, 25 searches after 5 warm-ups; GB/s = algorithmic bytes (rows x 2 x 384 x elem) / time.\n\n"
                    "| dtype | row-steps in flight (U) | CTAs/SM | scan ms (mean) | GB/s (mean) | GB/s (best) | frac of 8 TB/s |\n|---|---:|---:|---:|---:|---:|---:|\n")
            for r in recs:
                f.write(f"| {r['dtype']} | {r['unroll']} | {r['bps']} | {r['scan_ms_mean']:.4f} | {r['gbs_mean']:.0f} | {r['gbs_best']:.0f} | {r['gbs_mean']/8000:.3f} |\n")
        print("wrote", name)

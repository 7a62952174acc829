#!/usr/bin/env python3
"""Print headline metrics + the top stall-sampled SASS lines of an .ncu-rep (first kernel)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); H, U, r = rows[0], rows[1], rows[2]
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
          "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread"]:
    if k in H: print(f"{k} = {r[H.index(k)]} {U[H.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); H = rows[1]; data = rows[2:]; ci = {h: i for i, h in enumerate(H)}
tot = sum(int(x[ci["# Samples"]]) for x in data); print("total samples", tot)
cols = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(x[ci[h]]) for x in data) for h in cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for x in sorted(data, key=lambda x: -int(x[ci["# Samples"]]))[:ntop]:
    st = dict(sorted({h: int(x[ci[h]]) for h in cols if int(x[ci[h]]) > 0}.items(), key=lambda kv: -kv[1])[:2])
    print(f"{int(x[ci['# Samples']]):7d} {100*int(x[ci['# Samples']])/tot:5.1f}% exec={x[ci['Instructions Executed']]:>9s} {x[ci['Source']].strip()[:64]:64s} {st}")

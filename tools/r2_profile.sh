#!/bin/bash
# ncu evidence for round 2 (one GPU): launch lists of the driver's bench command for the headline and the
# tensor-core workload, and one --set full capture of the top kernel of each.  Each ncu pass runs only after
# the same command has exited 0 without ncu.  Outputs -> gpurun_out/ (summarised by tools/summarize_profiles.py).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
B1="python bench.py --gpus 1 --steps 5 --warmup 3 --reps 1 --no-cpu-baseline --no-dropin --secondary="
B2="python bench.py --gpus 1 --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --reps 1 --no-cpu-baseline --no-dropin --secondary="
B3="python bench.py --gpus 1 --workload 10m_fp32_q1_top10 --steps 3 --warmup 3 --reps 1 --no-cpu-baseline --no-dropin --secondary="
echo "== plain runs"
timeout 600 $B1 > gpurun_out/plain1.log 2>&1; echo "1m rc=$?"
timeout 600 $B2 > gpurun_out/plain2.log 2>&1; echo "gemm rc=$?"
timeout 600 $B3 > gpurun_out/plain3.log 2>&1; echo "10m rc=$?"
echo "== launch lists"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B1 > gpurun_out/ncu1.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 613 -c 80 --csv --log-file gpurun_out/launches_gemm.csv $B2 > gpurun_out/ncu3.log 2>&1; echo "rc=$?"
echo "== full captures"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cab_scan -s 6 -c 1 -o gpurun_out/prof_gemv -f $B1 > gpurun_out/ncu2.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_scan -s 3 -c 1 -o gpurun_out/prof_gemm -f $B2 > gpurun_out/ncu4.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cab_scan -s 4 -c 1 -o gpurun_out/prof_gemv10m -f $B3 > gpurun_out/ncu5.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 6 -c 1 -o gpurun_out/prof_finalize -f $B1 > gpurun_out/ncu6.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep

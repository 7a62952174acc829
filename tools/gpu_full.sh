#!/bin/bash
# Full single-GPU round: tests, all bench workloads, ncu launch lists + full captures. Logs -> gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches*.csv
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for wl in 1m_fp32_q1_top10 10m_fp32_q1_top10 10m_bf16_q1_top10 10m_bf16_q256_top100; do
  steps=200; [ "$wl" != "1m_fp32_q1_top10" ] && steps=50; [ "$wl" = "10m_bf16_q256_top100" ] && steps=40
  timeout 900 python bench.py --workload $wl --steps $steps --warmup 10 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); r=d['roofline']
print('  ms/step', round(d['ms_per_step'],4), 'value', round(d['value'],1), 'q/s', round(d['queries_per_s'],1), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'roof', round(r['achieved'],1), r['unit'], 'frac', round(r['frac'],3), d['clocks'], 'cpu', d.get('cpu_baseline',{}).get('value'))"
done
cp gpurun_out/bench_1m_fp32_q1_top10.json gpurun_out/bench.json
echo "== ncu launches (gemv 1M)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/plain1.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_scan -s 4 -c 2 -o gpurun_out/prof_gemv -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo "rc=$?"
echo "== ncu (gemm 10M)"
timeout 900 python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 620 -c 60 --csv --log-file gpurun_out/launches_gemm.csv python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1; echo "rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_scan -s 2 -c 1 -o gpurun_out/prof_gemm -f python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1; echo "rc=$?"

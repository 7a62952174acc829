#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('1M fp32: ms/step', d['ms_per_step'], 'q/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'scan GB/s', d['roofline']['achieved'])"
timeout 900 python bench.py --no-cpu-baseline --workload 10m_fp32_q1_top10 --steps 50 --warmup 5 > gpurun_out/bench_10m_fp32.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_10m_fp32.json')); print('10M fp32: ms/step', d['ms_per_step'], 'q/s', d['queries_per_s'], 'e2e ms', d['e2e']['ms_per_step'], 'scan GB/s', d['roofline']['achieved'], d['clocks'])"
timeout 600 python tools/sweep_gemv.py 1000000 fp32 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        r=json.loads(l); print(r['dtype'], r['unroll'], r['bps'], round(r['scan_ms_mean'],4), round(r['gbs_mean']))"

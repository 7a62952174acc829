#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json
echo "== bench gemm"; timeout 900 python bench.py --workload 10m_bf16_q256_top100 --steps 20 --warmup 3 > gpurun_out/bench_gemm.json 2> gpurun_out/bench_gemm.err; echo "rc=$?"; cat gpurun_out/bench_gemm.json
rm -f gpurun_out/prof_gemv.ncu-rep
timeout 900 python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gemm_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_scan -s 2 -c 1 -o gpurun_out/prof_gemm -f python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_gemm.csv python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_launch.log 2>&1
echo "ncu launches rc=$?"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -15 gpurun_out/pytest_gemm.log
timeout 900 python bench.py --workload 10m_bf16_q256_top100 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gemm.json 2> gpurun_out/bench_gemm.err; echo "bench gemm rc=$?"; cat gpurun_out/bench_gemm.json; tail -5 gpurun_out/bench_gemm.err
if [ "$1" = "ncu" ]; then
timeout 900 python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gemm_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_scan -s 2 -c 1 -o gpurun_out/prof_gemm -f python bench.py --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"
fi

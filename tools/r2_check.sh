#!/bin/bash
# Round 2 first GPU pass: smoke, GPU parity tests, short benches. Logs -> gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench 1m"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "rc=$?"; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
echo "== bench gemm"; timeout 600 python bench.py --workload 10m_bf16_q256_top100 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_gemm.json 2> gpurun_out/bench_gemm.err; echo "rc=$?"; cat gpurun_out/bench_gemm.json; tail -3 gpurun_out/bench_gemm.err

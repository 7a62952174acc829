#!/bin/bash
# Round 2 GPU pass: smoke, GPU parity tests, the driver's bench command. Logs -> gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench (driver command)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver.json 2> gpurun_out/bench_driver.err; echo "rc=$?"; cat gpurun_out/bench_driver.json; tail -5 gpurun_out/bench_driver.err

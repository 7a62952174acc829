#!/bin/bash
# One gpurun call: smoke, GPU parity tests, a short bench, sweeps, ncu launch list. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" = "sweep" ]; then
  echo "== sweep 1M"; timeout 600 python tools/sweep_gemv.py 1000000 > gpurun_out/sweep_1m.jsonl 2>&1; cat gpurun_out/sweep_1m.jsonl
  echo "== sweep 10M"; timeout 900 python tools/sweep_gemv.py 10000000 > gpurun_out/sweep_10m.jsonl 2>&1; cat gpurun_out/sweep_10m.jsonl
fi
if [ "$1" = "ncu" ] || [ "$2" = "ncu" ]; then
  echo "== ncu launches"
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_scan -s 4 -c 2 -o gpurun_out/prof_gemv -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi

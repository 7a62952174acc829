#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python tools/finalize_probe.py 2>&1 | grep -v Warn | tail -8
B2="python bench.py --gpus 1 --workload 10m_bf16_q256_top100 --steps 3 --warmup 3 --reps 1 --no-cpu-baseline --no-dropin --secondary="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 613 -c 80 --csv --log-file gpurun_out/launches_gemm.csv $B2 > gpurun_out/ncu3.log 2>&1; echo "gemm launch list rc=$?"
echo "== driver bench N=1"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_driver.json 2> gpurun_out/bench_driver.err; echo "rc=$?"; tail -3 gpurun_out/bench_driver.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>/dev/null; echo "rc=$?"; cut -c1-300 gpurun_out/bench_ref.json

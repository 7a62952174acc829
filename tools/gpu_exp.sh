#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -3
for thr in 10 0.1; do
timeout 600 python bench.py --workload 10m_bf16_q256_top100 --steps 10 --warmup 3 --no-cpu-baseline --threshold $thr 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('thr', d['config']['threshold'], 'ms/step', round(d['ms_per_step'],3), 'TF/s', round(d['roofline']['achieved'],1), 'hbm', round(d['roofline']['hbm_gbs'],0), 'e2e ms', round(d['e2e']['ms_per_step'],3))"
done

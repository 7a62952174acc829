#!/bin/bash
# Multi-GPU bench exactly as the driver launches it: torchrun, one rank per GPU.  $1 = N, $2.. = extra workloads
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() {  # workload steps port
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $N --steps $2 --warmup 10 --no-cpu-baseline --workload $1 > gpurun_out/bench_n${N}_$1.json 2> gpurun_out/bench_n${N}_$1.err
  echo "bench N=$N $1 rc=$?"; grep '^{' gpurun_out/bench_n${N}_$1.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  value', round(d['value'],1), 'q/s', round(d['queries_per_s'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'scan', round(r['achieved'],1), r['unit'], 'global segs', d['config']['global_segments'])"
  tail -2 gpurun_out/bench_n${N}_$1.err | cut -c1-300
}
run 1m_fp32_q1_top10 100 29511
p=29520
for wl in "$@"; do p=$((p+1)); run $wl 30 $p; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --impl reference 2>/dev/null | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('reference arm: value', round(d['value'],2), 'cores', d['cpu_baseline']['cores'])"

#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): two-device test, then the driver's scaling commands at every N <= visible GPUs.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "visible GPUs: $NG"
echo "== two-device / peer tests"; timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_multi.log
for N in ${NS:-1 2 4 8}; do
  [ $N -le $NG ] || continue
  echo "== bench N=$N"
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --secondary '' --no-dropin --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 $EXTRA > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "repetitions")}, d["e2e"]["ms_per_step"], d["parity_check"]["ok"], d.get("exchange_breakdown"))
    for s in d.get("secondary", []):
        print("  secondary", s["config"]["workload"], s["ms_per_step"], s["e2e"]["ms_per_step"], s["roofline"]["frac"], s["parity_check"]["ok"], s.get("exchange_breakdown", {}).get("max_over_ranks"))
except Exception as e:
    print("no line:", e)
PY
  tail -3 gpurun_out/scale_n$N.err
done

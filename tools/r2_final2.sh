#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
echo "== bench: fp32 tensor-core route + worst cases"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-dropin --secondary 10m_fp32_q256_top10_tc,10m_bf16_q256_top100,10m_bf16_q256_top100_ascending,10m_bf16_q256_top100_clustered,10m_fp32_q1_top10,10m_fp32_q1_top10_ascending,10m_fp32_q1_top10_clustered > gpurun_out/bench_worst.json 2> gpurun_out/bench_worst.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_worst.json").read().strip().splitlines()[-1])
print("primary", d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity_check"]["ok"])
for s in d["secondary"]:
    print(s["config"]["workload"], "ms", round(s["ms_per_step"], 4), "min", round(s["repetitions"]["ms_per_step_min"], 4), "e2e", round(s["e2e"]["ms_per_step"], 4), "kernel", round(s["roofline"]["kernel_ms"], 4), "frac", round(s["roofline"]["frac"], 3), s["clocks"]["sm_mhz"], s["parity_check"]["ok"], s["parity_check"]["problems"][:2], s.get("shadow"))
PY
tail -5 gpurun_out/bench_worst.err

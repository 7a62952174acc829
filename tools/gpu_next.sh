#!/bin/bash
# One gpurun call for the section 8(f) rank 3/4 rows: GPU tests (incl. the new ones), smoke,
# score_all timing, a short default bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_legacy.py tests/test_gpu_engine.py -m gpu -q > gpurun_out/pytest_new.log 2>&1; echo "pytest-new rc=$?"; tail -40 gpurun_out/pytest_new.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== score_all"; timeout 600 python tools/bench_score_all.py 50 > gpurun_out/score_all.jsonl 2> gpurun_out/score_all.err; echo "score_all rc=$?"; cat gpurun_out/score_all.jsonl; tail -5 gpurun_out/score_all.err
echo "== bench"; timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

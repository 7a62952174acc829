#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/recall_sweep.py 10000000 4096 > gpurun_out/recall_sweep.json 2> gpurun_out/recall_sweep.err; echo "rc=$?"; cat gpurun_out/recall_sweep.json; tail -3 gpurun_out/recall_sweep.err

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -40 gpurun_out/pytest_gemm.log

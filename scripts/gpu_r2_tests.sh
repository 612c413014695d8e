#!/bin/bash
# round 2: all GPU tests (incl. the N-rank ones when N GPUs are visible); logs kept for profiles/
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x ${PYTEST_ARGS} > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r2.log

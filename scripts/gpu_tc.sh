#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "tensor_core or full_size_against" > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -30 gpurun_out/pytest_tc.log
timeout 300 python scripts/spmm_probe.py eval > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; grep score_topk gpurun_out/probe_eval.log

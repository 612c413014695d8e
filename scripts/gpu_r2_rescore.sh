#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/eval_probe2.py > gpurun_out/eval_probe_rs32.jsonl 2> gpurun_out/eval_probe_rs32.err; echo "probe rc=$?"; cat gpurun_out/eval_probe_rs32.jsonl
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 200 -k "tensor_core or score_topk_tc or tc_" > gpurun_out/pytest_tc_rs32.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tc_rs32.log

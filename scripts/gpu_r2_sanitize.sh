#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_model.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_model.log
timeout 300 python scripts/sanitize_step.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
for tool in ${SAN_TOOLS}; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_step.py ${SAN_ARGS} > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize ok|========= (Invalid|Race|Error|Uninit)" gpurun_out/sanitizer_$tool.log | head -8
done

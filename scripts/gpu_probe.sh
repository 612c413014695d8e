#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python scripts/spmm_probe.py ${PROBE:-spmm} > gpurun_out/probe.log 2>&1; echo "probe rc=$?"
tail -3 gpurun_out/probe.log
if [ -n "$NCU" ]; then
timeout 300 python scripts/spmm_probe.py ncu > gpurun_out/ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -c 3 -o gpurun_out/spmm_r1b -f python scripts/spmm_probe.py ncu > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
fi

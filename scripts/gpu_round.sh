#!/bin/bash
# tests -> bench -> probe, one call
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${BENCH_STEPS:-500} --warmup 10 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
if [ -n "$PROBE" ]; then timeout 900 python scripts/spmm_probe.py $PROBE > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; fi

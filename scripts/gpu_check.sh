#!/bin/bash
# One gpurun call: smoke -> GPU tests -> short bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?" | tee -a gpurun_out/bench.err
tail -3 gpurun_out/bench.log ; tail -5 gpurun_out/bench.err

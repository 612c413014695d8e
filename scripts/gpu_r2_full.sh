#!/bin/bash
# driver-like: all GPU tests, smoke, reference arm, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/bench_ref.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.log

#!/bin/bash
# what the driver runs at round end
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --gpus 1 --steps 1500 --warmup 20 --no-cpu-baseline > gpurun_out/bench_long.log 2> gpurun_out/bench_long.err; echo "bench long rc=$?"; tail -1 gpurun_out/bench_long.log | cut -c1-400

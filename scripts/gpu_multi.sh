#!/bin/bash
# multi-GPU validation: NCCL tests, then bench at N ranks in both modes
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -5 gpurun_out/pytest_dist.log
for mode in dp rowpart; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${BENCH_STEPS:-300} --warmup 10 --parallel $mode > gpurun_out/bench_${mode}_$N.log 2> gpurun_out/bench_${mode}_$N.err; echo "bench $mode rc=$?"; tail -1 gpurun_out/bench_${mode}_$N.log; tail -3 gpurun_out/bench_${mode}_$N.err
done

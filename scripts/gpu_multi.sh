#!/bin/bash
# multi-GPU validation: NCCL tests, bench at N ranks in the three modes, scaled power-law rowpart
N=${NGPU:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 > gpurun_out/pytest_dist.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/pytest_dist.log
fi
for mode in ${MODES:-dp_idx dp rowpart}; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${BENCH_STEPS:-300} --warmup 10 --parallel $mode > gpurun_out/bench_${mode}_$N.log 2> gpurun_out/bench_${mode}_$N.err; echo "bench $mode rc=$?"; tail -1 gpurun_out/bench_${mode}_$N.log | cut -c1-700; tail -2 gpurun_out/bench_${mode}_$N.err
done
if [ -n "$POWERLAW" ]; then
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/powerlaw_rowpart.py $POWERLAW > gpurun_out/powerlaw_$N.log 2> gpurun_out/powerlaw_$N.err; echo "powerlaw rc=$?"; tail -1 gpurun_out/powerlaw_$N.log; tail -3 gpurun_out/powerlaw_$N.err
fi

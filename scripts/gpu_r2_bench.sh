#!/bin/bash
# round 2: bench at N GPUs (N from NGPU), reference arm, optional extras
N=${NGPU:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.log; tail -5 gpurun_out/bench_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${BENCH_STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
fi
if [ -n "$CEILING" ]; then timeout 300 python scripts/l2_gather_ceiling.py > gpurun_out/l2_gather_ceiling.jsonl 2> gpurun_out/l2_gather_ceiling.err; echo "ceiling rc=$?"; cat gpurun_out/l2_gather_ceiling.jsonl | cut -c1-400; fi
if [ -n "$REFARM" ]; then timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.log; fi

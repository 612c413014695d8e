#!/bin/bash
# round 2, final N-GPU pass: one multi-process feature-partition test, then the driver's command (plain bench.py --gpus N: headline under the
# feature partition, amazon-book, the row partition beside it, large_graph under the row partition with the one-GPU step on rank 0)
N=${NGPU:-2}
mkdir -p gpurun_out
T0=$(date +%s)
LGCN_TEST_RANKS=$N timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 500 -k "feature_partition_matches_reference and graph" > gpurun_out/pytest_feat_final_${N}ranks.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -3 gpurun_out/pytest_feat_final_${N}ranks.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_default_n$N.json 2> gpurun_out/bench_default_n$N.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s lines=$(wc -l < gpurun_out/bench_default_n$N.json)"; tail -c 1500 gpurun_out/bench_default_n$N.json; grep -v "Warn\|warn\|return torch\|^\*\|^$\|precision\|CudaIPC" gpurun_out/bench_default_n$N.err | tail -8

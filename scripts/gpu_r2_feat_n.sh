#!/bin/bash
# feature partition on N real GPUs: multi-process tests (model API, procedures), then bench at N (headline featpart, amazon-book, rowpart beside it)
N=${NGPU:-2}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
LGCN_TEST_RANKS=$N timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -k "${TESTS:-feature}" > gpurun_out/pytest_feat_${N}ranks.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_feat_${N}ranks.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_ARGS:---no-large-graph} > gpurun_out/bench_feat_n$N.json 2> gpurun_out/bench_feat_n$N.err; echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_feat_n$N.json; grep -v "Warn\|warn\|return torch\|^\*\|^$" gpurun_out/bench_feat_n$N.err | tail -8

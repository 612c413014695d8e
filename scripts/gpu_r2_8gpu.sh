#!/bin/bash
# round 2, one 8-GPU call: N-rank tests, bench at N=8 (incl. config 5 with the one-GPU step on rank 0), a 3x config-5 graph, latency probe
N=${NGPU:-8}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host_${N}.txt; nproc >> gpurun_out/host_${N}.txt; nvidia-smi -L >> gpurun_out/host_${N}.txt
LGCN_TEST_RANKS=$N timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 600 -k "rowpart-] or rowpart] or procedures or barrier" > gpurun_out/pytest_dist_${N}ranks.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dist_${N}ranks.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_n$N.log; grep -v "Warn\|warn\|return torch\|^\*" gpurun_out/bench_n$N.err | tail -5
timeout 300 $TR --master-port 29513 scripts/rowpart_probe.py > gpurun_out/rowpart_probe_$N.log 2> gpurun_out/rowpart_probe_$N.err; echo "probe rc=$?"; tail -1 gpurun_out/rowpart_probe_$N.log | cut -c1-1500
if [ -n "$BIG" ]; then
timeout 900 $TR --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --large-scale $BIG --large-skip-1gpu > gpurun_out/bench_n${N}_big.log 2> gpurun_out/bench_n${N}_big.err; echo "big rc=$?"; tail -c 1800 gpurun_out/bench_n${N}_big.log; grep -v "Warn\|warn\|return torch\|^\*" gpurun_out/bench_n${N}_big.err | tail -5
fi

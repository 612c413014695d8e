#!/bin/bash
# one `ncu --set full` capture of both tensor-core passes of a warm full-size evaluation call
mkdir -p gpurun_out
timeout 300 python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 2 -c 2 -o gpurun_out/score_tc -f python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu_full.log 2>&1
echo "rc=$?"

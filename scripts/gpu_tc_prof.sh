#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/evaltc_launches.csv python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu.log 2>&1
echo "rc=$?"
timeout 300 python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_tc_kernel -s 1 -c 1 -o gpurun_out/score_tc -f python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu_full.log 2>&1
echo "rc=$?"

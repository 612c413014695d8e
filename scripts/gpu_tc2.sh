#!/bin/bash
# TC eval path: parity tests, timing probe, then a per-launch time list of one full-size call (ncu, times only)
mkdir -p gpurun_out; rm -f gpurun_out/probe.jsonl
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -x -k "tensor_core or full_size_against" > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?"; tail -5 gpurun_out/pytest_tc.log
timeout 300 python scripts/spmm_probe.py eval > gpurun_out/probe_eval.log 2>&1; echo "probe rc=$?"; grep score_topk_tc gpurun_out/probe_eval.log
timeout 300 python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/evaltc_launches.csv python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu.log 2>&1
echo "ncu rc=$?"
if [ -n "$AUX_FULL" ]; then
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"tc_mark|tc_select|rescore" -s 3 -c 3 -o gpurun_out/tc_aux -f python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu_aux.log 2>&1
echo "aux ncu rc=$?"
fi

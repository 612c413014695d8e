#!/bin/bash
mkdir -p gpurun_out
export LGCN_TC_DEBUG_MMA_ONLY=1
timeout 300 python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max --clock-control none --cache-control none -k regex:score_tc --csv --log-file gpurun_out/evaltc_dbg.csv python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu.log 2>&1
echo "ncu rc=$?"; grep -v "^==" gpurun_out/evaltc_dbg.csv | cut -d, -f5,13,15- | tail -12

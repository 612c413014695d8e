#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "blocking or blocked or row_block or rank_metrics" > gpurun_out/pytest_blocked.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_blocked.log
timeout 900 python scripts/blocked_probe.py --scale ${SCALE:-1.0} > gpurun_out/blocked_probe.jsonl 2> gpurun_out/blocked_probe.err; echo "probe rc=$?"; cat gpurun_out/blocked_probe.jsonl; tail -3 gpurun_out/blocked_probe.err
if [ -n "$NCU" ]; then
timeout 900 ncu -k regex:spmm_kernel --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none --csv --log-file gpurun_out/blocked_ncu.csv python scripts/blocked_probe.py --scale ${SCALE:-1.0} --ncu > gpurun_out/blocked_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/blocked_ncu.log; wc -l gpurun_out/blocked_ncu.csv
fi

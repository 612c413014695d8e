#!/bin/bash
# feature partition, 1-GPU pass: kernel tests at the slice widths, emulated-rank tests, what a rank's step costs at each width
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --timeout 300 -k "feature_partition or spmm_vs_oracle or bpr_vs_oracle or spmm_adam_epilogue" > gpurun_out/pytest_feat_emulated.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_feat_emulated.log
timeout 600 python scripts/feat_probe.py gowalla amazon-book > gpurun_out/feat_probe.jsonl 2> gpurun_out/feat_probe.err; echo "probe rc=$?"; cat gpurun_out/feat_probe.jsonl; tail -3 gpurun_out/feat_probe.err

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/feat_probe.py gowalla amazon-book > gpurun_out/feat_probe2.jsonl 2> gpurun_out/feat_probe2.err; echo "probe rc=$?"; cat gpurun_out/feat_probe2.jsonl; tail -3 gpurun_out/feat_probe2.err

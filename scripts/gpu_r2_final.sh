#!/bin/bash
# round 2, final 1-GPU pass: what the driver runs (tests, smoke, reference arm, default bench) + ncu --set full of the ranking kernels
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - T0 ))s"; tail -2 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - T0 ))s"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$? t=$(( $(date +%s) - T0 ))s lines=$(wc -l < gpurun_out/bench_ref.json)"; cut -c1-300 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s lines=$(wc -l < gpurun_out/bench_n1.json)"; cut -c1-600 gpurun_out/bench_n1.json
timeout 300 python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"score_tc_kernel|tc_select_kernel|rescore_kernel" -s 4 -c 4 -o gpurun_out/r2_score_tc_full -f python scripts/spmm_probe.py evaltc > gpurun_out/evaltc_ncu_full.log 2>&1
echo "ncu rc=$? t=$(( $(date +%s) - T0 ))s"
ncu -i gpurun_out/r2_score_tc_full.ncu-rep --page raw --csv > gpurun_out/r2_score_tc_ncu_raw.csv 2>/dev/null; wc -l gpurun_out/r2_score_tc_ncu_raw.csv
rm -f gpurun_out/r2_score_tc_full.ncu-rep

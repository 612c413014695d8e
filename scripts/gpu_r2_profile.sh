#!/bin/bash
# round 2 profiling: bench (plain), launch list of the same command, ncu --set full of K1 inside it
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --no-extra --no-large-graph --no-baselines"
timeout 600 python bench.py $ARGS > gpurun_out/bench_prof_plain.log 2> gpurun_out/bench_prof_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r2_bench_launches_ncu.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_kernel -s 40 -c 8 -o gpurun_out/r2_spmm_gowalla_full -f python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r2_spmm_gowalla_full.ncu-rep --page raw --csv > gpurun_out/r2_spmm_gowalla_ncu_raw.csv 2>/dev/null; wc -l gpurun_out/r2_spmm_gowalla_ncu_raw.csv
rm -f gpurun_out/r2_spmm_gowalla_full.ncu-rep

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/feat_ncu_target.py > gpurun_out/feat_ncu_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_kernel|spmm_narrow_kernel" -c 8 -o gpurun_out/r2_k1_slices_full -f python scripts/feat_ncu_target.py > gpurun_out/feat_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/r2_k1_slices_full.ncu-rep --page raw --csv > gpurun_out/r2_k1_slice_widths_ncu_raw.csv 2>/dev/null; wc -l gpurun_out/r2_k1_slice_widths_ncu_raw.csv
rm -f gpurun_out/r2_k1_slices_full.ncu-rep

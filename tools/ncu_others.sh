#!/bin/bash
# --set full captures of the non-dominant kernels of a config-3 solve (first Newton iteration)
tag=${1:-v11}
for k in k_backsolve_fused k_stage_fused k_ls_eval k_lambda_step_w; do
  skip=0; [ $k = k_backsolve_fused ] && skip=14   # the bottom (widest) level is the last launch of the sweep
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/${tag}_$k -f \
      python tools/profile_solve.py --depth 16 --max-iter 1 > gpurun_out/${tag}_$k.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_$k.txt 2>/dev/null
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_${k}_src.csv 2>/dev/null
  echo "=== $k"; grep -E "^\s+(void )?k_|Duration|DRAM Throughput|L1/TEX Cache Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Memory Throughput  " gpurun_out/${tag}_$k.txt | head -9
done

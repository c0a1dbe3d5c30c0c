#!/bin/bash
# --set full captures of the kernels added in the second half of round 2 (config 3): k_step_affine (2nd launch = with the
# stage update), k_unpack_run / k_gather_run (DMA path: needs the caller buffers page-locked, tools/e2e_probe.py 16 pin)
tag=${1:-v19}
ncu --set full --clock-control none --import-source on -k regex:k_step_affine -s 1 -c 1 -o gpurun_out/${tag}_k_step_affine -f \
    python tools/profile_solve.py --depth 16 --max-iter 4 > gpurun_out/${tag}_k_step_affine.log 2>&1
for k in k_unpack_run k_gather_run; do
  ncu --set full --clock-control none -k regex:$k -s 0 -c 1 -o gpurun_out/${tag}_$k -f python tools/e2e_probe.py 16 pin 1 > gpurun_out/${tag}_$k.log 2>&1
done
for k in k_step_affine k_unpack_run k_gather_run; do
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_$k.txt 2>/dev/null
  echo "=== $k"; grep -E "^\s+(void )?k_|Duration|DRAM Throughput|L1/TEX Cache Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Memory Throughput  " gpurun_out/${tag}_$k.txt | head -9
done

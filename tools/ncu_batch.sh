#!/bin/bash
# ncu evidence for the config-4 batch kernels (chains, paired chain parents): one --set full capture per
# kernel + a DRAM-bytes/duration pass over one batch solve.  Outputs under gpurun_out/<tag>_*.
tag=${1:-v9b}
N=${2:-4096}
for k in k_build_factor_chain2 k_backsolve_chain2 k_stage_fused; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/${tag}_$k -f \
      python tools/batch_probe.py --n $N --reps 1 --profile 0 --e2e 0 --ref 0 > gpurun_out/${tag}_$k.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_$k.txt 2>/dev/null
done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/batch_probe.py --n $N --reps 1 --profile 0 --e2e 0 --ref 0 > gpurun_out/${tag}_launches.log 2>&1
grep -c "k_" gpurun_out/${tag}_launches.csv

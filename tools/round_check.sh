#!/bin/bash
# Round trip used for the record: GPU parity tests, headline bench (defaults) + reference arm, one ncu pass for
# the launch list and the DRAM traffic per launch of every kernel class (tools/ncu_traffic.py turns it into
# profiles/traffic_latest.json, which bench.py reports as roofline.traffic), and --set full captures of the two
# level kernels.
tag=${1:-dev}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --impl reference > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -3 gpurun_out/bench_$tag.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_traffic.csv python tools/profile_solve.py --depth 16 --max-iter 2 > gpurun_out/${tag}_traffic.log 2>&1
python tools/ncu_traffic.py gpurun_out/${tag}_traffic.csv 16 gpurun_out/${tag}_traffic.json "ncu dram__bytes_{read,write}.sum, 2 Newton iterations of the depth-16 QP ($tag)" | cut -c1-700
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("ms/solve", round(d["ms_per_step"], 3), "e2e", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["e2e"].items()})
print("clocks", d["clocks"], "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"]["ok"])
print({k: (round(x["ms"], 2), x["launches"]) for k, x in d["roofline"]["per_kernel"].items()})
for k, v in d["other_workloads"].items():
    print(k, json.dumps(v)[:700])
PY
# --set full captures: bottom level of the leaves-to-root sweep, bottom level of the root-to-leaves sweep
ncu --set full --clock-control none --import-source on -k regex:k_build_factor -c 1 -o gpurun_out/${tag}_bf -f \
    python tools/profile_solve.py --depth 16 --max-iter 1 > gpurun_out/${tag}_bf.log 2>&1
ncu -i gpurun_out/${tag}_bf.ncu-rep --page details > gpurun_out/${tag}_ncu_bf.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_backsolve_tma -s 1 -c 1 -o gpurun_out/${tag}_bs -f \
    python tools/profile_solve.py --depth 16 --max-iter 1 > gpurun_out/${tag}_bs.log 2>&1
ncu -i gpurun_out/${tag}_bs.ncu-rep --page details > gpurun_out/${tag}_ncu_bs_tma.txt 2>/dev/null
for f in gpurun_out/${tag}_ncu_bf.txt gpurun_out/${tag}_ncu_bs_tma.txt; do
  echo "== $f"; grep -E "^\s+(void )?k_|Duration|DRAM Throughput|L1/TEX Cache Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy|Memory Throughput  " $f | head -9
done
rm -f gpurun_out/${tag}_bf.ncu-rep gpurun_out/${tag}_bs.ncu-rep

#!/bin/bash
# A/B of build variants: DEPTH=16 tools/ab_bench.sh <variant-dir-under-treeqp_b200/lib> ...   ("." = default build)
DEPTH=${DEPTH:-16}
for var in "$@"; do
  lib=treeqp_b200/lib/$var/libtreeqp_b200.so
  TREEQP_B200_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --depth $DEPTH --no-cpu-baseline --no-extra > gpurun_out/ab_${var}_d$DEPTH.json 2> gpurun_out/ab_${var}_d$DEPTH.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_${var}_d$DEPTH.json"))
print("$var depth $DEPTH", "ms/solve", round(d["ms_per_step"], 3), {k: round(x["ms"], 2) for k, x in d["roofline"]["per_kernel"].items() if x["ms"] > 0.02 * d["ms_per_step"]})
PY
done

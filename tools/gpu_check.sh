#!/bin/bash
# GPU round trip used during development: parity tests, headline bench, latency-floor bench (depth 10)
tag=${1:-dev}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --steps 5 --warmup 3 --depth 10 --no-cpu-baseline --no-extra > gpurun_out/bench_${tag}_d10.json 2>> gpurun_out/bench_$tag.err
python - <<PY
import json
for f in ("gpurun_out/bench_$tag.json", "gpurun_out/bench_${tag}_d10.json"):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "ms/solve", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3))
    print("  ", {k: (round(x["ms"], 2), x["launches"]) for k, x in d["roofline"]["per_kernel"].items()})
PY

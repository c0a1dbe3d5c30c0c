#!/usr/bin/env python
"""Summarise one ncu pass (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--csv) over a few Newton iterations into DRAM traffic per launch for every kernel class of bench.py.

    python tools/ncu_traffic.py gpurun_out/v8_traffic.csv 16 profiles/traffic_latest.json "ncu ... (r01 v8)"
"""
import csv
import json
import re
import sys

src, depth, dst = sys.argv[1], int(sys.argv[2]), sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else src

CLASS = [(r"k_build_factor", "k_build_factor"), (r"k_backsolve_fused|k_backsolve_chain|k_backsolve_tma", "k_backsolve_fused"),
         (r"k_stage_fused<.*, 0>", "k_stage_fused<0>"), (r"k_stage_fused<.*, 2>", "k_stage_fused<2>"),
         (r"k_stage_fused<.*, 3>", "k_stage_fused<3>"), (r"k_ls_eval", "k_ls_eval"),
         (r"k_stage<0>", "k_stage<0>"), (r"k_stage<2>", "k_stage<2>"), (r"k_trial", "k_trial"),
         (r"k_factor_level", "k_factor_level"), (r"k_backsolve_level", "k_backsolve_level"), (r"k_hessian", "k_hessian"),
         (r"k_lambda_step", "k_lambda_step")]

rows = list(csv.reader(open(src, errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mn, mu, mv, idc = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "second": 1e6, "s": 1e6}
launch = {}
for r in rows[h + 1:]:
    if len(r) <= mv:
        continue
    d = launch.setdefault(r[idc], {"name": r[kn]})
    d[r[mn]] = float(r[mv].replace(",", "")) * SCALE.get(r[mu], 1.0)
acc = {}
for d in launch.values():
    cls = next((c for pat, c in CLASS if re.search(pat, d["name"])), None)
    if cls is None:
        continue
    a = acc.setdefault(cls, {"launches": 0, "bytes": 0.0, "us": 0.0})
    a["launches"] += 1
    a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
out = {"depth": depth, "source": note,
       "per_launch_bytes": {k: round(v["bytes"] / v["launches"]) for k, v in acc.items()},
       "per_launch_us_under_ncu": {k: round(v["us"] / v["launches"], 2) for k, v in acc.items()},
       "launches_seen": {k: v["launches"] for k, v in acc.items()}}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out))

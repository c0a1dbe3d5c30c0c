#!/usr/bin/env python
"""one-screen summary of a bench.py JSON line:  python tools/summarize_bench.py gpurun_out/bench_x.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r, e = d["roofline"], d["e2e"]
print("value %.2f %s  %.3f ms/step  %.0f us/Newton iteration  | e2e %.2f (%.2f ms: pack %.2f h2d %.2f gpu %.2f d2h %.2f)" % (
    d["value"], d["unit"], d["ms_per_step"], d["us_per_newton_iter"], e["value"], e["ms_per_step"], e["pack_ms"], e["h2d_ms"], e["gpu_ms"], e["d2h_ms"]))
print("roofline %s: %.0f of %.0f GB/s = %.3f, traffic/launch %s vs algorithmic %.0f" % (
    r["kernel"], r["achieved"], r["peak"], r["frac"], r["traffic"], r["algorithmic_bytes_per_launch"]))
print("clocks", d["clocks"], " launches", d["gpu_launches"], " kkt", d["config"].get("kkt"), d["config"].get("kkt_device"))
if d.get("cpu_baseline"):
    c = d["cpu_baseline"]
    print("cpu %.3f %s on %d cores (%s) -> %.0fx resident, %.0fx e2e" % (c["value"], c["unit"], c["cores"], c["kind"], d["value"] / c["value"], e["value"] / c["value"]))
print({k: (x["ms"], x["launches"]) for k, x in r["per_kernel"].items()})
for k, v in (d.get("other_workloads") or {}).items():
    print(" ", k, json.dumps(v)[:600])
if d.get("sharded_single_tree"):
    print("  sharded", d["sharded_single_tree"])

#!/bin/bash
# per-launch durations (ncu, serialised and cold-cache: compare levels with each other, not with the bench) of the
# leaves-to-root sweep of Newton iteration `it` (1-based) of a config-3 solve: shows what a level costs when its blocks
# are re-used (iteration >= 3) against a full factorisation (TREEQP_B200_NO_SKIP=1)
tag=${1:-lv}; iters=${2:-4}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
    python tools/profile_solve.py --depth 16 --max-iter $iters > gpurun_out/${tag}_launches.log 2>&1
python - <<PY
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/${tag}_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); im = hdr.index("Metric Name"); iv = hdr.index("Metric Value"); ii = hdr.index("ID")
launch = collections.OrderedDict()
for r in rows[1:]:
    launch.setdefault(r[ii], {"k": r[ik]})[r[im]] = float(r[iv].replace(",", ""))
seq = list(launch.values())
# split into Newton iterations at every k_stage_fused<..., 0> launch
it, out = 0, collections.defaultdict(list)
for l in seq:
    if re.search(r"k_stage_fused<\d+, \d+, 0>", l["k"]): it += 1
    out[it].append(l)
for it in sorted(out):
    bf = [l for l in out[it] if "k_build_factor" in l["k"]]
    if not bf: continue
    t = [l["gpu__time_duration.sum"] / 1e3 for l in bf]
    by = [(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) / 1e6 for l in bf]
    st = [l["gpu__time_duration.sum"] / 1e3 for l in out[it] if "k_stage_fused" in l["k"]]
    bs = sum(l["gpu__time_duration.sum"] / 1e3 for l in out[it] if "k_backsolve" in l["k"])
    print("iteration %d: k_build_factor sum %.1f us, widest levels (bottom first) us %s, MB %s | k_stage_fused %s | backsolve sweep %.1f" % (
        it, sum(t), [round(x, 1) for x in t[:7]], [round(x, 1) for x in by[:7]], [round(x, 1) for x in st], bs))
PY

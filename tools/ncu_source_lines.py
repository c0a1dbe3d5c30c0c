#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line:
share of warp-stall samples, instructions per block, top stall reasons.
    python tools/ncu_source_lines.py export.csv [nblocks] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
nblk = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
fname, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == '': continue
    try: ln = int(r[0]); s = int(r[hdr.index('# Samples')]); n = int(r[hdr.index('Instructions Executed')])
    except ValueError: continue
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try:
                if int(r[i]) > 0: st[h[6:]] = int(r[i])
            except ValueError: pass
    lines.append((fname, ln, r[1].strip()[:84], s, n, st))
tot = sum(l[3] for l in lines); tinst = sum(l[4] for l in lines)
print('total samples', tot, ' warp instructions per block', tinst / nblk)
for f, ln, src, s, n, st in sorted(lines, key=lambda x: -x[3])[:top]:
    t = sorted(st.items(), key=lambda x: -x[1])[:4]
    print(f"{f[:18]:18s}:{ln:4d} {100*s/max(tot,1):5.1f}% inst/blk {n/nblk:7.1f} | {src} | {t}")

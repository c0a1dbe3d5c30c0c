#!/usr/bin/env python
"""Phase stamps of k_build_factor (build with EXTRA=-DTQB_BF_TIMING into treeqp_b200/lib/timing):
    TREEQP_B200_LIB=$PWD/treeqp_b200/lib/timing/libtreeqp_b200.so python tools/bf_timing.py [--depth 16]
Prints, per level (slot 0 = root, slot s = levels with 2^(s-1) CTAs), the cycles / ns thread 0 of CTA 0
spent in: descriptor load, build (loads + DMMA), PDL wait, Schur loads, Cholesky, write-out, Schur product."""
import argparse, ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treeqp_b200 import capi, problems as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=16)
ap.add_argument("--max-iter", type=int, default=3)
args = ap.parse_args()
lib = capi.TreeQpLib()
prob = P.random_tree(P.binary_tree_nk(args.depth), 12, 4, seed=0, ubound=0.2)
qp_in = capi.QpIn(lib, prob)
solver = capi.Tdunes(lib, qp_in, maxIter=args.max_iter)
solver.set_dual_initialization(np.zeros(prob.n_lambda))
assert solver.upload() == capi.TREEQP_OK
for _ in range(2):
    solver.solve_resident()
buf = (ctypes.c_longlong * (32 * 16))()
assert lib.lib.tqb_debug_bf_timing(buf) == 0
t = np.array(buf[:]).reshape(32, 16)
names = ["start->desc", "build", "pdl wait", "schur loads", "cholesky", "L/Ut out", "schur+exit"]
print("slot  " + "  ".join("%12s" % n for n in names) + "   | total cycles, total ns, end(ns) rel. to root end")
end_root = t[0, 15]
for s in range(32):
    if t[s, 7] == 0:
        continue
    c = t[s, :8]; g = t[s, 8:]
    print("%4d  " % s + "  ".join("%12d" % (c[i + 1] - c[i]) for i in range(7)) + "   | %d %d %d" % (c[7] - c[0], g[7] - g[0], g[7] - end_root))
    print("  ns  " + "  ".join("%12d" % (g[i + 1] - g[i]) for i in range(7)))

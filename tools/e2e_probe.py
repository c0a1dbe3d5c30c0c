"""phase times of treeqp_tdunes_solve on config 3 (TREEQP_B200_VERBOSE=1 prints them from the C side)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from treeqp_b200 import capi, problems as P
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lib = capi.TreeQpLib(capi.PRODUCT_LIB)
prob = P.random_tree(P.binary_tree_nk(depth), 12, 4, seed=0, ubound=0.2)
qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
if len(sys.argv) > 2 and sys.argv[2] == "pin":
    print("page-locked:", qp_in.pin() and qp_out.pin())
s = capi.Tdunes(lib, qp_in, maxIter=200)
for i in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    s.set_dual_initialization(np.zeros(prob.n_lambda))
    t0 = time.perf_counter(); st = s.solve(qp_out); t = time.perf_counter() - t0
    print("solve %d: status %d, %.3f ms wall, stats %s" % (i, st, 1e3 * t, {k: round(v, 3) for k, v in s.stats().items() if k in ("pack_ms", "h2d_ms", "gpu_ms", "d2h_ms")}), flush=True)

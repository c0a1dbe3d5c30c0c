#!/usr/bin/env python
"""Minimal driver for ncu: build the benchmark QP, upload once, run `--solves` resident solves.
    ncu ... python tools/profile_solve.py --depth 16 --solves 1 [--max-iter 2]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treeqp_b200 import capi, problems as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=16)
ap.add_argument("--solves", type=int, default=1)
ap.add_argument("--max-iter", type=int, default=200)
ap.add_argument("--workload", default="random")
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
lib = capi.TreeQpLib()
if args.workload == "random":
    prob = P.random_tree(P.binary_tree_nk(args.depth), 12, 4, seed=0, ubound=0.2)
else:
    prob = P.spring_mass(3, 2, 10, xmax1=0.2)
qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
solver = capi.Tdunes(lib, qp_in, maxIter=args.max_iter)
solver.set_dual_initialization(np.zeros(prob.n_lambda))
assert solver.upload() == capi.TREEQP_OK
for _ in range(args.solves):
    st = solver.solve_resident()
s = solver.stats()
print("status", st, "iter", s["iter"], "gpu_ms", s["gpu_ms"], "launches", s["kernel_launches"])

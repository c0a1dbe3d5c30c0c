#!/usr/bin/env python
"""Minimal driver for ncu on the config-5 trees (dense stage QPs, generic level kernels): build one of bench_dense's
problems, upload once, run `--solves` resident solves.
    ncu --metrics gpu__time_duration.sum ... python tools/dense_probe.py --problem linear_chain --solves 1"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treeqp_b200 import capi, problems as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problem", default="linear_chain", choices=["linear_chain", "quadcopter", "crane"])
ap.add_argument("--solves", type=int, default=1)
ap.add_argument("--max-iter", type=int, default=200)
args = ap.parse_args()
lib = capi.TreeQpLib()
if args.problem == "linear_chain":
    prob = P.linear_chain(8, 7, 4, 2, 20, seed=2)
    prob.umin[:] = -0.3; prob.umax[:] = 0.3; prob.xmin[:] = -1e12; prob.xmax[:] = 1e12
    prob = P.densify(prob, 6, 0.1, 0.0)
elif args.problem == "quadcopter":
    prob = P.densify_scaled(P.quadcopter(3, 2, 20, i_mpc=14, x0=[0.03, -0.03, 0.01, 0.8, -0.8, 0.3]), 4)
else:
    prob = P.densify_scaled(P.crane(3, 2, 20, x0=[0.0, 0.2, 0.1, -0.3]), 5)
qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
solver = capi.Tdunes(lib, qp_in, qp_solver=1, maxIter=args.max_iter)
solver.set_dual_initialization(np.zeros(prob.n_lambda))
assert solver.upload() == capi.TREEQP_OK
best = 1e9
for _ in range(args.solves):
    st = solver.solve_resident()
    best = min(best, solver.stats()["gpu_ms"])
s = solver.stats()
print("problem", args.problem, "status", st, "iter", s["iter"], "gpu_ms best", best, "launches", s["kernel_launches"])

"""A/B probe of the persistent small-tree kernel (tdunes_small.cuh) against the level-synchronous driver on single
small trees and on the config-4 batch: resident solve time (best of --reps), launches, iteration counts, and the
unmodified reference on one host core where oracle/_ref travelled.  TREEQP_B200_SMALL is read at create."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_port as OP
import cases as CS
from treeqp_b200 import capi, problems as P

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--ref", type=int, default=1)
ap.add_argument("--warps", default="")
args = ap.parse_args()
lib = capi.TreeQpLib(capi.PRODUCT_LIB)
ref = capi.TreeQpLib(OP.REF_LIB) if (args.ref and os.path.exists(OP.REF_LIB)) else None


def timed(prob, small, opts, warps=None):
    os.environ["TREEQP_B200_SMALL"] = "1" if small else "0"
    if warps: os.environ["TREEQP_B200_SMALL_WARPS"] = str(warps)
    else: os.environ.pop("TREEQP_B200_SMALL_WARPS", None)
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    s = capi.Tdunes(lib, qp_in, **opts)
    s.set_dual_initialization(np.zeros(prob.n_lambda))
    assert s.upload() == capi.TREEQP_OK
    best, wall = 1e9, 1e9
    for r in range(args.reps):
        t0 = time.perf_counter(); st = s.solve_resident(); w = time.perf_counter() - t0
        best = min(best, s.stats()["gpu_ms"]); wall = min(wall, 1e3 * w)
    s.download(qp_out)
    out = (best, wall, st, qp_out.iter, s.stats()["kernel_launches"], qp_out.x.copy())
    s.destroy()
    return out


single = [("spring_mass_cfg1", P.spring_mass(3, 2, 10), {}),
          ("spring_mass_tight", P.spring_mass(3, 2, 10, xmax1=0.2), {}),
          ("spring_mass_cfg2tree", P.spring_mass(3, 4, 20), {}),
          ("linear_chain_cfg4", P.linear_chain(4, 3, 3, 2, 20, seed=0), {}),
          ("quadcopter_cfg5", P.quadcopter(3, 2, 20, i_mpc=14, x0=CS.QUAD_X0), dict(maxIter=200)),
          ("crane_cfg5", P.crane(3, 2, 20, x0=CS.CRANE_X0), dict(maxIter=200)),
          ("random_d6 (127 nodes, nx 12)", P.random_tree(P.binary_tree_nk(6), 12, 4, seed=1, ubound=0.5), dict(maxIter=200)),
          ("random_d8 (511 nodes)", P.random_tree(P.binary_tree_nk(8), 12, 4, seed=3, ubound=0.5), dict(maxIter=200)),
          ("random_d10 (2047 nodes)", P.random_tree(P.binary_tree_nk(10), 12, 4, seed=2, ubound=0.2), dict(maxIter=200)),
          ("random_d12 (8191 nodes)", P.random_tree(P.binary_tree_nk(12), 12, 4, seed=2, ubound=0.2), dict(maxIter=200))]
print("%-32s %6s | %-28s | %-28s | %s" % ("single tree", "nodes", "level-synchronous", "persistent CTA", "reference, 1 core"))
for name, prob, opts in single:
    a = timed(prob, 0, opts)
    b = timed(prob, 1, opts)
    extra = ""
    for w in [int(x) for x in args.warps.split(",") if x]:
        c = timed(prob, 1, opts, w)
        extra += "  w%d %.3f" % (w, c[0])
    rt = ""
    if ref is not None and prob.Nn <= 3000:
        rin, rout = capi.QpIn(ref, prob), capi.QpOut(ref, prob)
        rs = capi.Tdunes(ref, rin, **opts)
        tt = 1e9
        for r in range(5):
            rs.set_dual_initialization(np.zeros(prob.n_lambda))
            rs.solve(rout)
            tt = min(tt, rout.total_time)
        rt = "%.3f ms, %d it" % (1e3 * tt, rout.iter)
    same = "x equal %.1e" % float(np.max(np.abs(a[5] - b[5])))
    print("%-32s %6d | %7.3f ms (wall %7.3f) %3d it %5d l | %7.3f ms (wall %7.3f) %3d it %3d l | %s | %s%s" % (
        name, prob.Nn, a[0], a[1], a[3], a[4], b[0], b[1], b[3], b[4], rt, same, extra), flush=True)

if args.batch > 0:
    probs = P.linear_chain_batch(args.batch, seed=0)
    qp_ins = [capi.QpIn(lib, p) for p in probs]
    qp_outs = [capi.QpOut(lib, p) for p in probs]
    res = {}
    for small, warps in ((0, None), (1, None), (1, 8)):
        os.environ["TREEQP_B200_SMALL"] = "1" if small else "0"
        if warps: os.environ["TREEQP_B200_SMALL_WARPS"] = str(warps)
        else: os.environ.pop("TREEQP_B200_SMALL_WARPS", None)
        batch = capi.TdunesBatch(lib, qp_ins)
        assert batch.upload() == capi.TREEQP_OK
        best = 1e9
        for r in range(max(3, args.reps // 2)):
            batch.solve_resident()
            best = min(best, batch.stats()["gpu_ms"])
        batch.download(qp_outs)
        iters = np.array([o.iter for o in qp_outs])
        xs = np.concatenate([o.x.ravel() for o in qp_outs[:64]])
        res[(small, warps)] = (iters, xs)
        print("batch %d x linear_chain, %s%s: %.3f ms per batch = %.0f QP/s, launches %d, converged %d, iterations max %d" % (
            args.batch, "persistent CTA" if small else "level-synchronous", " (%d warps)" % warps if warps else "", best,
            1e3 * args.batch / best, batch.stats()["kernel_launches"], int((batch.status == 0).sum()), iters.max()), flush=True)
        batch.destroy()
    print("iteration counts identical:", bool(np.array_equal(res[(0, None)][0], res[(1, None)][0])),
          " max |x_small - x_level| on 64 QPs: %.2e" % float(np.max(np.abs(res[(0, None)][1] - res[(1, None)][1]))),
          " 4 vs 8 warps bit-identical:", bool(np.array_equal(res[(1, None)][1], res[(1, 8)][1])))

"""timing probe for config 4: n independent linear_chain tree QPs (nm=4, nu=3, md=3, Nr=2, Nh=20 -> 175 nodes)
solved as one forest; compares with the compiled reference on a sample of the same QPs"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_port as OP
from treeqp_b200 import capi, problems as P

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--profile", type=int, default=1)
ap.add_argument("--e2e", type=int, default=2)
ap.add_argument("--ref", type=int, default=1)
args = ap.parse_args()
lib = capi.TreeQpLib(capi.PRODUCT_LIB)
t0 = time.perf_counter()
probs = P.linear_chain_batch(args.n, seed=0)
qp_ins = [capi.QpIn(lib, p) for p in probs]
qp_outs = [capi.QpOut(lib, p) for p in probs]
print("built %d QPs in %.1f s" % (args.n, time.perf_counter() - t0), flush=True)
batch = capi.TdunesBatch(lib, qp_ins)
assert batch.upload() == capi.TREEQP_OK
for r in range(args.reps):
    t0 = time.perf_counter(); st = batch.solve_resident(); t = time.perf_counter() - t0
    s = batch.stats()
    print("resident: %.2f ms wall, gpu %.2f ms, max iter %d, launches %d, %.0f QP/s" % (1e3 * t, s["gpu_ms"], s["iter"], s["kernel_launches"], args.n / t), flush=True)
assert (batch.status == 0).all(), np.unique(batch.status, return_counts=True)
for r in range(args.e2e):
    t0 = time.perf_counter(); st = batch.solve(qp_outs); t = time.perf_counter() - t0
    s = batch.stats()
    print("e2e: %.2f ms wall (pack %.2f, h2d %.2f, gpu %.2f, d2h %.2f) %.0f QP/s" % (1e3 * t, s["pack_ms"], s["h2d_ms"], s["gpu_ms"], s["d2h_ms"], args.n / t), flush=True)
if args.e2e == 0:
    batch.solve(qp_outs) if False else None
iters = np.array([o.iter for o in qp_outs])
print("iterations: min %d median %d max %d" % (iters.min(), np.median(iters), iters.max()))
if args.profile:
    lib.lib.treeqp_b200_set_profiling(1)
    batch.solve_resident()
    s = batch.stats()
    lib.lib.treeqp_b200_set_profiling(0)
    for k, v in sorted(s["kernels"].items(), key=lambda kv: -kv[1][0]):
        print("   %-22s %8.3f ms %6d launches" % (k, v[0], v[1]))
if args.ref and args.e2e and os.path.exists(OP.REF_LIB):
    ref = capi.TreeQpLib(OP.REF_LIB)
    m = min(args.n, 64)
    tt = 0.0
    for i in range(m):
        rin, rout = capi.QpIn(ref, probs[i]), capi.QpOut(ref, probs[i])
        rs = capi.Tdunes(ref, rin)
        rs.set_dual_initialization(np.zeros(probs[i].n_lambda))
        rs.solve(rout)
        assert rout.iter == iters[i], (i, rout.iter, iters[i])
        tt += rout.total_time
    print("reference CPU (1 thread): %.3f ms per QP on the first %d QPs -> %.0f QP/s per core; iteration counts identical" % (1e3 * tt / m, m, m / tt))

"""timing probe: SDUNES config 2 (spring-mass md=3 Nr=4 Nh=20) and friends, product vs compiled reference"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as CS
import oracle_port as OP
from treeqp_b200 import capi

lib = capi.TreeQpLib(capi.PRODUCT_LIB)
ref = capi.TreeQpLib(OP.REF_LIB) if os.path.exists(OP.REF_LIB) else None
for name in ["spring_mass_cfg1", "spring_mass_cfg2", "spring_mass_cfg2_u01", "linear_chain_u05", "linear_chain_nm8_md4"]:
    prob, opts = CS.get_sdunes_case(name)
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    s = capi.Sdunes(lib, qp_in, **opts)
    z = (np.zeros(s.n_lambda), np.zeros(s.n_mu))
    s.set_dual_initialization(*z)
    s.upload()
    best, e2e = 1e9, 1e9
    for r in range(10):
        s.solve_resident(); st = s.stats(); best = min(best, st["gpu_ms"])
    for r in range(10):
        s.set_dual_initialization(*z)
        t0 = time.perf_counter(); status = s.solve(qp_out); e2e = min(e2e, 1e3 * (time.perf_counter() - t0))
    line = "%-24s Ns=%3d it=%2d status=%d  gpu %.3f ms  e2e %.3f ms" % (name, s.Ns, qp_out.iter, status, best, e2e)
    if ref is not None:
        rin, rout = capi.QpIn(ref, prob), capi.QpOut(ref, prob)
        rs = capi.Sdunes(ref, rin, **opts)
        tb = 1e9
        for r in range(10):
            rs.set_dual_initialization(*z); rs.solve(rout); tb = min(tb, 1e3 * rout.total_time)
        line += "  | reference CPU %.3f ms (it=%d)" % (tb, rout.iter)
    print(line, flush=True)
    print("      phases (ms): " + "  ".join("%s %.3f" % (k.replace("sd_", ""), v[0]) for k, v in st["kernels"].items()), flush=True)

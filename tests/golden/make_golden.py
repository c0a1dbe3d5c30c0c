#!/usr/bin/env python
"""Regenerates the committed fixtures in tests/golden/ from the reference.

Runs ONLY in the build container (needs /root/reference and oracle/_ref/libtreeqp_ref.so built
by `make -C oracle ref`).  The GPU box never runs this: tests read the .npz files.

  spring_mass_data.npz   numeric arrays parsed out of examples/spring_mass_utils/data.c and
                         x0.txt (the reference's own example data; lambda0_tree.txt is all zeros)
  sdunes_<case>.npz      the same for treeqp_sdunes_solve (x0 eliminated)
  tdunes_<case>.npz      outputs of the UNMODIFIED reference (treeqp_tdunes_solve, clipping) on the
                         seeded problems of treeqp_b200/problems.py: status, Newton iterations,
                         x, u, lam, mu_x, mu_u, KKT residual.  Inputs are regenerated from the
                         seed by the tests, only a checksum of them is stored.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
REF = os.environ.get("TREEQP_REFERENCE", "/root/reference")


def parse_c_arrays(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"^(int|double)\s+(\w+)(\[\])?\s*=\s*(\{[^}]*\}|[^;]+);", txt, re.M):
        typ, name, arr, val = m.groups()
        if arr:
            nums = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", val)]
            out[name] = np.asarray(nums, dtype=np.float64 if typ == "double" else np.int64)
        else:
            out[name] = int(val) if typ == "int" else float(val)
    return out


def make_spring_mass_data():
    d = parse_c_arrays(os.path.join(REF, "examples/spring_mass_utils/data.c"))
    x0 = np.loadtxt(os.path.join(REF, "examples/spring_mass_utils/x0.txt"))
    lam0 = np.loadtxt(os.path.join(REF, "examples/spring_mass_utils/lambda0_tree.txt"))
    assert np.all(lam0 == 0.0)
    keep = ["A", "B", "b", "dR", "dQ", "dP", "r", "q", "p", "umin", "umax", "xmin", "xmax"]
    np.savez(os.path.join(HERE, "spring_mass_data.npz"), NX=d["NX"], NU=d["NU"], md=d["md"],
             Nr=d["Nr"], Nh=d["Nh"], x0=x0, **{k: d[k] for k in keep})
    print("spring_mass_data.npz: NX=%d NU=%d realisations=%d" % (d["NX"], d["NU"], len(d["A"]) // d["NX"] ** 2))


from cases import cases, checksum, dense_cases, sdunes_cases  # noqa: E402  (tests/cases.py)
import cases as CS  # noqa: E402

ONLY = set(os.environ.get("GOLDEN_ONLY", "").split(",")) - {""}   # GOLDEN_ONLY=name1,name2: just these cases


def make_tdunes_goldens():
    from treeqp_b200 import capi
    ref = capi.TreeQpLib(os.path.join(ROOT, "oracle/_ref/libtreeqp_ref.so"))
    for name, prob, opts in cases():
        if ONLY and name not in ONLY:
            continue
        sol = capi.solve_tdunes(ref, prob, **opts)
        np.savez_compressed(os.path.join(HERE, "tdunes_%s.npz" % name), status=sol.status, iter=sol.iter,
                            x=sol.x, u=sol.u, lam=sol.lam, mu_x=sol.mu_x, mu_u=sol.mu_u, kkt=sol.kkt,
                            input_sha256=checksum(prob), opts=repr(opts))
        print("%-28s Nn=%6d status=%d iter=%3d kkt=%.2e time=%.3f ms" % (
            name, prob.Nn, sol.status, sol.iter, sol.kkt, 1e3 * sol.total_time))


def make_dense_goldens():
    """outputs of the UNMODIFIED reference with qp_solver = TREEQP_QPOASES_SOLVER (qpOASES_e stage QPs)"""
    from treeqp_b200 import capi
    ref = capi.TreeQpLib(os.path.join(ROOT, "oracle/_ref/libtreeqp_ref.so"))
    for name, prob, opts in dense_cases():
        if ONLY and name not in ONLY:
            continue
        sol = capi.solve_tdunes(ref, prob, **opts)
        np.savez_compressed(os.path.join(HERE, "tdunes_dense_%s.npz" % name), status=sol.status, iter=sol.iter,
                            x=sol.x, u=sol.u, lam=sol.lam, mu_x=sol.mu_x, mu_u=sol.mu_u, kkt=sol.kkt,
                            input_sha256=checksum(prob), opts=repr(opts))
        print("dense  %-28s Nn=%6d status=%d iter=%3d kkt=%.2e time=%.3f ms" % (
            name, prob.Nn, sol.status, sol.iter, sol.kkt, 1e3 * sol.total_time))


def make_sdunes_goldens():
    """outputs of the UNMODIFIED reference treeqp_sdunes_solve (dual_Newton_scenarios.c:1850)"""
    from treeqp_b200 import capi
    ref = capi.TreeQpLib(os.path.join(ROOT, "oracle/_ref/libtreeqp_ref.so"))
    for name, prob, opts in sdunes_cases():
        if ONLY and name not in ONLY:
            continue
        sol = capi.solve_sdunes(ref, prob, **opts)
        np.savez_compressed(os.path.join(HERE, "sdunes_%s.npz" % name), status=sol.status, iter=sol.iter,
                            x=sol.x, u=sol.u, lam=sol.lam, mu_x=sol.mu_x, mu_u=sol.mu_u, kkt=sol.kkt,
                            input_sha256=checksum(prob), opts=repr(opts))
        print("sdunes %-28s Nn=%6d status=%d iter=%3d kkt=%.2e time=%.3f ms" % (
            name, prob.Nn, sol.status, sol.iter, sol.kkt, 1e3 * sol.total_time))


def make_general_goldens():
    """outputs of the UNMODIFIED reference with general stage constraints (QProblem path of qpOASES_e)"""
    from treeqp_b200 import capi
    ref = capi.TreeQpLib(os.path.join(ROOT, "oracle/_ref/libtreeqp_ref.so"))
    for name, prob, opts in CS.general_cases():
        if ONLY and name not in ONLY:
            continue
        qp_in, qp_out = capi.QpIn(ref, prob), capi.QpOut(ref, prob)
        s = capi.Tdunes(ref, qp_in, **opts)
        s.set_dual_initialization(np.zeros(prob.n_lambda))
        st = s.solve(qp_out)
        kkt = float(ref.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
        assert st == 0 and kkt < 1e-8, (name, st, kkt)
        np.savez_compressed(os.path.join(HERE, "tdunes_general_%s.npz" % name), status=st, iter=qp_out.iter, x=qp_out.x, u=qp_out.u,
                            lam=qp_out.lam, mu_x=qp_out.mu_x, mu_u=qp_out.mu_u, mu_d=qp_out.mu_d, kkt=kkt,
                            input_sha256=checksum(prob), opts=repr(opts))
        print("general %-28s Nn=%6d rows=%4d status=%d iter=%3d kkt=%.2e active rows %d" % (
            name, prob.Nn, int(prob.nc.sum()), st, qp_out.iter, kkt, int((qp_out.mu_d != 0).sum())))


def make_large_golden():
    """config 3's QP at depth 14 through the UNMODIFIED reference (OpenMP build: 32 767 nodes)"""
    from treeqp_b200 import capi
    ref = capi.TreeQpLib(os.path.join(ROOT, "oracle/_ref/libtreeqp_ref_omp.so"))
    prob, opts = CS.large_case()
    sol = capi.solve_tdunes(ref, prob, **opts)
    d = CS.large_summary(prob, dict(status=sol.status, iter=sol.iter, x=sol.x, u=sol.u, lam=sol.lam, mu_x=sol.mu_x, mu_u=sol.mu_u))
    np.savez_compressed(os.path.join(HERE, "tdunes_large_%s.npz" % CS.LARGE_NAME), kkt=sol.kkt, input_sha256=checksum(prob),
                        opts=repr(opts), stride=CS.LARGE_STRIDE, **d)
    print("large  %-28s Nn=%6d status=%d iter=%3d kkt=%.2e time=%.1f s, active inputs %d" % (
        CS.LARGE_NAME, prob.Nn, sol.status, sol.iter, sol.kkt, sol.total_time,
        int(np.unpackbits(d["u_lower"]).sum() + np.unpackbits(d["u_upper"]).sum())))


if __name__ == "__main__":
    what = sys.argv[1:] or ["data", "tdunes", "sdunes", "dense", "large", "general"]
    if "general" in what:
        make_general_goldens()
    if "large" in what:
        make_large_golden()
    if "dense" in what:
        make_dense_goldens()
    if "data" in what:
        make_spring_mass_data()
    if "tdunes" in what:
        make_tdunes_goldens()
    if "sdunes" in what:
        make_sdunes_goldens()

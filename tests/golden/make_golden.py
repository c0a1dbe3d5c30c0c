#!/usr/bin/env python
"""Regenerates the committed fixtures in tests/golden/ from the reference.

Runs ONLY in the build container (needs /root/reference and oracle/_ref/libtreeqp_ref.so built
by `make -C oracle ref`).  The GPU box never runs this: tests read the .npz files.

  spring_mass_data.npz   numeric arrays parsed out of examples/spring_mass_utils/data.c and
                         x0.txt (the reference's own example data; lambda0_tree.txt is all zeros)
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


def checksum(prob):
    import hashlib
    h = hashlib.sha256()
    for name in ["nk", "nx", "nu", "A", "B", "b", "Qd", "Rd", "q", "r", "xmin", "xmax", "umin", "umax"]:
        h.update(np.ascontiguousarray(getattr(prob, name)).tobytes())
    return h.hexdigest()


def cases():
    from treeqp_b200 import problems as P
    yield "spring_mass_cfg1", P.spring_mass(3, 2, 10), {}
    yield "spring_mass_tight", P.spring_mass(3, 2, 10, xmax1=0.2), {}
    yield "spring_mass_cfg2tree", P.spring_mass(3, 4, 20), {}
    yield "random_d6_u05", P.random_tree(P.binary_tree_nk(6), 12, 4, seed=1, ubound=0.5), dict(maxIter=200)
    yield "random_d10_u02", P.random_tree(P.binary_tree_nk(10), 12, 4, seed=2, ubound=0.2), dict(maxIter=200)
    yield "random_d8_u05", P.random_tree(P.binary_tree_nk(8), 12, 4, seed=3, ubound=0.5), dict(maxIter=200)
    yield "random_vary_dims", P.random_tree(P.multistage_nk(3, 2, 5), 5, 3, seed=4, ubound=0.3, vary_dims=True), dict(maxIter=200)
    yield "random_ternary_chain", P.random_tree(P.multistage_nk(3, 2, 8), 6, 2, seed=5, ubound=0.2, xbound=3.0), dict(maxIter=200)
    yield "linear_chain_cfg4", P.linear_chain(4, 3, 3, 2, 20, seed=0), {}
    yield "linear_chain_cfg4_s7", P.linear_chain(4, 3, 3, 2, 20, seed=7), {}
    yield "linear_chain_pinned_x0", P.linear_chain(2, 1, 2, 2, 6, seed=1, eliminate_x0=False), {}


def make_tdunes_goldens():
    from treeqp_b200 import capi
    ref = capi.TreeQpLib(os.path.join(ROOT, "oracle/_ref/libtreeqp_ref.so"))
    for name, prob, opts in cases():
        sol = capi.solve_tdunes(ref, prob, **opts)
        np.savez_compressed(os.path.join(HERE, "tdunes_%s.npz" % name), status=sol.status, iter=sol.iter,
                            x=sol.x, u=sol.u, lam=sol.lam, mu_x=sol.mu_x, mu_u=sol.mu_u, kkt=sol.kkt,
                            input_sha256=checksum(prob), opts=repr(opts))
        print("%-28s Nn=%6d status=%d iter=%3d kkt=%.2e time=%.3f ms" % (
            name, prob.Nn, sol.status, sol.iter, sol.kkt, 1e3 * sol.total_time))


if __name__ == "__main__":
    make_spring_mass_data()
    make_tdunes_goldens()

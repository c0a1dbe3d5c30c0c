"""Would honouring checkLastActiveSet (dual_Newton_tree.c:334-371, 563, 703) pay on the GPU?

The reference skips the re-factorisation of the blocks behind `idxFactorStart`: the suffix of the parent numbering
whose own, kids' and descendants' active sets did not change since the last Newton iteration.  A GPU version could
skip any block (not only a suffix) whose whole subtree is unchanged: its L, Utilde and Schur complement are those of
the last iteration, only the right-hand side parts have to be redone.

This script measures the opportunity on config 3: it solves the QP with maxIter = 1, 2, ... (the iterates are
deterministic, so run k reproduces the first k iterations), reads the stage solutions back and counts, per Newton
iteration, (a) blocks whose own + kids' active sets are unchanged, (b) blocks whose whole subtree is unchanged
(= what may be skipped), (c) the reference's suffix rule."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from treeqp_b200 import capi, problems as P

ap = argparse.ArgumentParser()
ap.add_argument("--depth", type=int, default=16)
args = ap.parse_args()
lib = capi.TreeQpLib(capi.PRODUCT_LIB)
prob = P.random_tree(P.binary_tree_nk(args.depth), 12, 4, seed=0, ubound=0.2)
Nn = prob.Nn
dad = prob.dad
nk = np.asarray(prob.nk)
parents = np.nonzero(nk > 0)[0]
qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
full = capi.solve_tdunes(lib, prob, maxIter=200)
print("config 3 at depth %d: %d nodes, %d parent blocks, %d Newton iterations" % (args.depth, Nn, parents.size, full.iter))


def active(k):
    s = capi.Tdunes(lib, qp_in, maxIter=k)
    s.set_dual_initialization(np.zeros(prob.n_lambda))
    s.solve(qp_out)
    x, u = qp_out.x, qp_out.u
    s.destroy()
    ax = (x == prob.xmax).astype(np.int8) - (x == prob.xmin).astype(np.int8)
    au = (u == prob.umax).astype(np.int8) - (u == prob.umin).astype(np.int8)
    return ax, au


def per_node_any(flags, dims):
    """flags: one bool per element of the concatenated per-node vectors -> one bool per node"""
    out = np.zeros(Nn, dtype=bool)
    node_of = np.repeat(np.arange(Nn), dims)
    np.logical_or.at(out, node_of, flags)
    return out


prev = active(1)
print("%4s %10s %12s %14s %16s" % ("iter", "nodes chg", "blocks same", "subtree same", "reference suffix"))
for k in range(2, full.iter + 1):
    cur = active(k)
    node_chg = per_node_any(cur[0] != prev[0], prob.nx) | per_node_any(cur[1] != prev[1], prob.nu)  # own active set changed
    blk_chg = node_chg.copy()                                    # block of parent p: p itself or one of its kids
    np.logical_or.at(blk_chg, dad[1:], node_chg[1:])
    sub_chg = blk_chg.copy()                                     # ... or any block below
    for g in range(Nn - 1, 0, -1):
        if sub_chg[g]: sub_chg[dad[g]] = True
    same_blk = int((~blk_chg[parents]).sum()); same_sub = int((~sub_chg[parents]).sum())
    chg_idx = np.nonzero(blk_chg[parents])[0]
    suffix = parents.size - (chg_idx.max() + 1 if chg_idx.size else 0)
    print("%4d %10d %11.1f%% %13.1f%% %15.1f%%" % (k, int(node_chg.sum()), 100.0 * same_blk / parents.size,
                                                   100.0 * same_sub / parents.size, 100.0 * suffix / parents.size), flush=True)
    prev = cur

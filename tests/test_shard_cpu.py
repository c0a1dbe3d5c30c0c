"""CPU tests of the subtree-sharding plan (host logic of the multi-GPU path, treeqp_b200_shard_plan):
run here without a GPU, the N>1 consistency check runs as two gloo processes."""
import os
import subprocess
import sys

import numpy as np
import pytest

from treeqp_b200 import capi, problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("depth,nranks,cut", [(6, 2, 1), (6, 4, 2), (8, 8, 3), (5, 1, 1)])
def test_plan_partitions_the_tree(depth, nranks, cut, product_lib):
    prob = P.random_tree(P.binary_tree_nk(depth), 12, 4, seed=0)
    qp_in = capi.QpIn(product_lib, prob)
    plans = [capi.shard_plan(product_lib, qp_in, r, nranks) for r in range(nranks)]
    owners = np.sum([p["owned"].astype(int) for p in plans], axis=0)
    assert (owners == 1).all(), "every node must be owned by exactly one rank"
    _, stage, _ = P.tree_tables(prob.nk)
    for r, p in enumerate(plans):
        assert p["cut_depth"] == cut and p["ncut"] == 2 ** cut
        assert p["ghost"].sum() == p["ncut"] - p["ncut"] // nranks
        assert (stage[p["ghost"]] == cut).all()
        # the levels above the cut belong to rank 0, every subtree to one rank
        assert p["owned"][stage < cut].all() == (r == 0)
        n_sub = (2 ** (depth + 1 - cut) - 1) * (p["ncut"] // nranks)
        assert p["owned"][stage >= cut].sum() == n_sub
        assert p["n_local"] == (2 ** (cut + 1) - 1) + n_sub - p["ncut"] // nranks


def test_plan_rejects_chains(product_lib):
    prob = P.spring_mass(1, 0, 5) if hasattr(P, "spring_mass") else None
    qp_in = capi.QpIn(product_lib, prob)
    with pytest.raises(ValueError):
        capi.shard_plan(product_lib, qp_in, 0, 2)


def test_plans_agree_across_two_gloo_processes():
    """world_size 2 over gloo: each process plans its own rank, then they exchange ownership masks"""
    worker = os.path.join(ROOT, "tests", "shard_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "partition ok" in outs[0]

"""gloo worker of tests/test_shard_cpu.py: rank r plans its shard, the ranks all-reduce ownership"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from treeqp_b200 import capi, problems as P  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", rank=rank, world_size=world)
lib = capi.TreeQpLib(capi.PRODUCT_LIB)
prob = P.random_tree(P.binary_tree_nk(7), 12, 4, seed=2)
plan = capi.shard_plan(lib, capi.QpIn(lib, prob), rank, world)
own = torch.from_numpy(plan["owned"].astype(np.int64))
dist.all_reduce(own)
cuts = torch.tensor([plan["cut_depth"], plan["ncut"]])
lo, hi = cuts.clone(), cuts.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert bool((own == 1).all()) and bool((lo == hi).all())
if rank == 0:
    print("partition ok: cut depth %d, %d cut nodes" % (plan["cut_depth"], plan["ncut"]))
dist.destroy_process_group()

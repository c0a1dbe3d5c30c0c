"""Worker of the multi-GPU parity test (launched by torch.distributed.run, one rank per GPU):
solves ONE tree QP sharded by subtree over the ranks and checks, on rank 0, that the assembled
solution equals the single-GPU solve of the same QP (same status, iteration count, trial counts;
x,u,lambda,mu to 1e-10)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as CS  # noqa: E402
from treeqp_b200 import capi, problems as P  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.TreeQpLib(capi.PRODUCT_LIB)
    assert lib.lib.treeqp_b200_set_device(local) == 0

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    comm = capi.Comm(lib, rank, world, local, bcast)
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    failures = []
    for name, prob, opts in [("random_d%d" % depth, P.random_tree(P.binary_tree_nk(depth), 12, 4, seed=3, ubound=0.2), dict(maxIter=200)),
                             ("random_d7_u05", P.random_tree(P.binary_tree_nk(7), 12, 4, seed=1, ubound=0.5), dict(maxIter=200)),
                             # infinity-norm termination: the ranks' partial norms ride on the cut exchange and are combined by max
                             ("random_d8_maxnorm", P.random_tree(P.binary_tree_nk(8), 12, 4, seed=3, ubound=0.5), dict(maxIter=200, termCondition=2)),
                             ("random_d8_2norm", P.random_tree(P.binary_tree_nk(8), 12, 4, seed=3, ubound=0.5), dict(maxIter=200, termCondition=1)),
                             # large enough for the DMA path: caller buffers page-locked, every rank moves its subtrees'
                             # levels as strided copies from / into them (and wide levels re-use unchanged blocks)
                             ("random_d14_pinned", P.random_tree(P.binary_tree_nk(14), 12, 4, seed=5, ubound=0.2), dict(maxIter=200))]:
        qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
        if name.endswith("pinned"):
            assert qp_in.pin() and qp_out.pin()
        sh = capi.TdunesSharded(lib, qp_in, comm, **opts)
        sh.set_dual_initialization(np.zeros(prob.n_lambda))
        st = sh.solve(qp_out)
        stats, plan = sh.stats(), sh.plan()
        own = capi.shard_plan(lib, qp_in, rank, world)["owned"]
        # assemble on every rank: owned parts are disjoint, so a SUM of masked vectors is the solution
        nxm = np.repeat(own, prob.nx); num = np.repeat(own, prob.nu); lamm = np.repeat(own[1:], prob.nx[1:])
        parts = {}
        for key, mask in (("x", nxm), ("u", num), ("lam", lamm), ("mu_x", nxm), ("mu_u", num)):
            t = torch.from_numpy(np.where(mask, getattr(qp_out, key), 0.0)).cuda()
            dist.all_reduce(t)
            parts[key] = t.cpu().numpy()
        cnt = torch.tensor(own.astype(np.int64)).cuda()
        dist.all_reduce(cnt)
        warm_ok = True
        if name.endswith("pinned"):
            # a new dual start after a DMA-mode upload travels alone (the pinned staging block does not hold the other
            # inputs of the nodes that came straight from the caller's buffer): restart at the optimum -> 0 iterations
            sh.set_dual_initialization(parts["lam"])
            st_w = sh.solve_resident()
            warm_ok = (st_w == 0 and sh.stats()["iter"] == 0)
        if rank == 0:
            ref = capi.solve_tdunes(lib, prob, **opts)
            ok = (st == ref.status and qp_out.iter == ref.iter and stats["ls_iters"] == ref.stats["ls_iters"]
                  and bool((cnt.cpu().numpy() == 1).all()))
            errs = {k: CS.rel_err(parts[k], getattr(ref, k)) for k in parts}
            ok = ok and all(e < 1e-10 for e in errs.values()) and warm_ok
            if name.endswith("pinned"):   # the DMA path moves the caller's rows: its byte count differs from the packed block's
                ok = ok and stats["h2d_bytes"] != 8 * 0 and stats["pack_ms"] >= 0.0
            print("%s: sharded over %d GPUs (cut depth %d, %d local nodes): status %d iter %d (single GPU: %d %d) gpu %.2f ms vs %.2f ms, max rel err %.1e -> %s"
                  % (name, world, plan["cut_depth"], plan["n_local"], st, qp_out.iter, ref.status, ref.iter,
                     stats["gpu_ms"], ref.stats["gpu_ms"], max(errs.values()), "OK" if ok else "MISMATCH"), flush=True)
            if not ok:
                failures.append(name)
        sh.destroy()
    comm.destroy()
    dist.barrier()
    dist.destroy_process_group()
    if failures:
        sys.exit(1)


if __name__ == "__main__":
    main()

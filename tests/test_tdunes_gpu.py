"""GPU parity tests: treeqp_tdunes_solve of libtreeqp_b200.so (CUDA) through the C ABI versus
(1) goldens produced by the unmodified reference, (2) the oracle port on the same inputs,
(3) the compiled reference itself when oracle/_ref travelled to the box.

Bar (BASELINE.json north_star): identical status, Newton iteration count, line-search trial
counts and active sets; x, u, lambda, mu within 1e-10 relative (FP64)."""
import numpy as np
import pytest

import cases as CS
import oracle_port as OP
from treeqp_b200 import capi, problems as P

pytestmark = pytest.mark.gpu
TOL = 1e-10


def check_against(sol, ref, prob, tol=TOL):
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(getattr(sol, k), ref[k]) < tol, k
    for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
        assert np.array_equal(getattr(sol, k) == lo, ref[k] == lo), "active set (lower) differs: " + k
        assert np.array_equal(getattr(sol, k) == hi, ref[k] == hi), "active set (upper) differs: " + k


def set_path(monkeypatch, path):
    """path=fast: compile-time-dimension fused kernels where the tree is uniform (tdunes_fused.cuh), one launch per
    phase and depth; path=small: the same device functions inside ONE persistent kernel, one CTA per tree
    (tdunes_small.cuh; the default for uniform trees of up to 1024 nodes, forced here whatever the size);
    path=generic: runtime-dimension kernels forced through TREEQP_B200_GENERIC=1 (all read at create)"""
    monkeypatch.setenv("TREEQP_B200_GENERIC", "1" if path == "generic" else "0")
    monkeypatch.setenv("TREEQP_B200_SMALL", "1" if path == "small" else "0")


PATHS = ["fast", "small", "generic"]
# uniform trees: the persistent kernel must be what ran (k_init + the kernel + copy-out: a handful of launches)
UNIFORM = {"spring_mass_cfg1", "spring_mass_tight", "spring_mass_cfg2tree", "random_d6_u05", "random_d10_u02", "random_d8_u05",
           "linear_chain_cfg4", "linear_chain_cfg4_s7", "quadcopter_cfg5", "crane_cfg5"}


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", CS.CASE_NAMES)
def test_solve_matches_reference_golden(name, path, product_lib, monkeypatch):
    set_path(monkeypatch, path)
    prob, opts = CS.get_case(name)
    g = CS.golden(name)
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert sol.status == int(g["status"])
    assert sol.iter == int(g["iter"]), "Newton iteration count differs from the reference"
    check_against(sol, g, prob)
    assert sol.kkt < 1e-8          # the reference's own acceptance threshold (spring_mass_dual_newton_tree.c:156)
    assert sol.stats["kernel_launches"] > 0
    if path == "small" and name in UNIFORM:
        assert sol.stats["kernel_launches"] <= 8, "the persistent small-tree kernel did not run"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name", ["spring_mass_tight", "random_d6_u05", "random_d8_u05", "random_d10_u02"])
def test_line_search_counts_match_oracle(name, path, product_lib, monkeypatch):
    set_path(monkeypatch, path)
    prob, opts = CS.get_case(name)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert sol.iter == s["iter"]
    assert sol.stats["ls_iters"] == s["ls_iters"], "per-iteration line-search trial counts differ"
    assert sol.stats["n_reg"] == s["n_reg"], "number of regularised blocks differs"


@pytest.mark.parametrize("opts", [dict(regType=1), dict(termCondition=0, stationarityTolerance=1e-14),
                                  dict(termCondition=1), dict(lineSearchBeta=0.5, lineSearchGamma=0.2), dict(lineSearchBeta=0.8, lineSearchGamma=0.05),
                                  dict(maxIter=7), dict(lineSearchMaxIter=3, maxIter=20)])
@pytest.mark.parametrize("path", ["fast", "small"])
def test_options_match_oracle(opts, path, product_lib, monkeypatch):
    set_path(monkeypatch, path)
    prob, _ = CS.get_case("spring_mass_tight")
    s = OP.solve_port(prob, **CS.port_opts(opts))
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert (sol.status, sol.iter) == (s["status"], s["iter"])
    # runs cut off by maxIter stop in the middle of an ill-conditioned path (regularised singular
    # blocks): mid-trajectory iterates agree to ~1e-6, converged solutions to 1e-9
    check_against(sol, s, prob, tol=1e-9 if sol.status == 0 else 1e-5)


def test_no_regularisation_on_regular_problem(product_lib):
    prob, _ = CS.get_case("spring_mass_cfg1")
    s = OP.solve_port(prob, regType=0)
    sol = capi.solve_tdunes(product_lib, prob, regType=0)
    assert (sol.status, sol.iter) == (s["status"], s["iter"]) == (0, 3)
    check_against(sol, s, prob)


def test_live_reference_when_available(product_lib, ref_lib):
    prob = P.random_tree(P.binary_tree_nk(7), 12, 4, seed=11, ubound=0.3)
    r = capi.solve_tdunes(ref_lib, prob, maxIter=200)
    s = capi.solve_tdunes(product_lib, prob, maxIter=200)
    assert (s.status, s.iter) == (r.status, r.iter)
    check_against(s, dict(x=r.x, u=r.u, lam=r.lam, mu_x=r.mu_x, mu_u=r.mu_u), prob)


def test_invalid_option_and_not_descent(product_lib):
    prob, _ = CS.get_case("spring_mass_cfg1")
    qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
    solver = capi.Tdunes(product_lib, qp_in, termCondition=7)
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    assert solver.solve(qp_out) == capi.TREEQP_INVALID_OPTION
    solver.destroy()


def test_warm_start_and_resolve(product_lib):
    """second solve from the optimal duals terminates at iteration 0 with the same solution
    (the closed-loop MPC pattern: set_dual_initialization(previous lam), solve)"""
    prob, _ = CS.get_case("spring_mass_tight")
    qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
    solver = capi.Tdunes(product_lib, qp_in)
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    assert solver.solve(qp_out) == 0
    x1, lam1, it1 = qp_out.x, qp_out.lam, qp_out.iter
    solver.set_dual_initialization(lam1)
    assert solver.solve(qp_out) == 0
    assert qp_out.iter == 0 and it1 == 58
    assert CS.rel_err(qp_out.x, x1) < 1e-12
    # cold start again: bit-identical to the first run (deterministic reductions)
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    assert solver.solve(qp_out) == 0
    assert qp_out.iter == it1 and np.array_equal(qp_out.x, x1) and np.array_equal(qp_out.lam, lam1)
    solver.destroy()


def test_large_tree_properties(product_lib):
    """Size-independent checks at a size the oracle would need minutes for: KKT of the reference's
    checker, dual feasibility of the line search, repeatability."""
    prob = P.random_tree(P.binary_tree_nk(13), 12, 4, seed=5, ubound=0.2)
    a = capi.solve_tdunes(product_lib, prob, maxIter=200)
    b = capi.solve_tdunes(product_lib, prob, maxIter=200)
    assert a.status == 0 and a.kkt < 1e-8
    assert (a.iter, a.stats["ls_iters"]) == (b.iter, b.stats["ls_iters"])
    assert np.array_equal(a.x, b.x) and np.array_equal(a.lam, b.lam)
    assert np.all(a.u <= prob.umax) and np.all(a.u >= prob.umin)
    assert np.all(a.mu_u[a.u == prob.umax] >= 0) and np.all(a.mu_u[a.u == prob.umin] <= 0)


def test_batch_matches_individual_solves(product_lib):
    """treeqp_tdunes_solve_batch (config-4 shape: independent linear_chain MPC QPs that differ in
    x0) returns, per QP, exactly what treeqp_tdunes_solve returns for that QP alone"""
    probs = P.linear_chain_batch(12, seed=3)
    singles = [capi.solve_tdunes(product_lib, p) for p in probs]
    qp_ins = [capi.QpIn(product_lib, p) for p in probs]
    qp_outs = [capi.QpOut(product_lib, p) for p in probs]
    batch = capi.TdunesBatch(product_lib, qp_ins)
    assert batch.solve(qp_outs) == capi.TREEQP_OK
    assert len(set(s.iter for s in singles)) > 1, "the instances should need different iteration counts"
    for i, (s, o) in enumerate(zip(singles, qp_outs)):
        assert batch.status[i] == s.status == 0
        assert o.iter == s.iter
        assert np.array_equal(o.x, s.x) and np.array_equal(o.u, s.u) and np.array_equal(o.lam, s.lam), i
        assert np.array_equal(o.mu_x, s.mu_x) and np.array_equal(o.mu_u, s.mu_u), i
    g = CS.golden("linear_chain_cfg4")   # seed 0 instance against the reference golden
    p0 = P.linear_chain_batch(1, seed=0, x0_kick=1.5, ubound=2.0, xbound=2.0)[0]
    qi, qo = capi.QpIn(product_lib, p0), capi.QpOut(product_lib, p0)
    b1 = capi.TdunesBatch(product_lib, [qi])
    assert b1.solve([qo]) == capi.TREEQP_OK and qo.iter == int(g["iter"])
    assert CS.rel_err(qo.x, g["x"]) < TOL and CS.rel_err(qo.lam, g["lam"]) < TOL
    batch.destroy(); b1.destroy()


def test_batch_mixed_shapes_and_warm_start(product_lib):
    """a forest may mix tree shapes; per-QP dual starts are honoured"""
    probs = [CS.get_case("spring_mass_cfg1")[0], CS.get_case("random_vary_dims")[0], CS.get_case("linear_chain_cfg4")[0]]
    singles = [capi.solve_tdunes(product_lib, p, maxIter=200) for p in probs]
    qp_ins = [capi.QpIn(product_lib, p) for p in probs]
    qp_outs = [capi.QpOut(product_lib, p) for p in probs]
    batch = capi.TdunesBatch(product_lib, qp_ins, maxIter=200)
    assert batch.solve(qp_outs) == capi.TREEQP_OK
    for s, o in zip(singles, qp_outs):
        assert o.iter == s.iter and CS.rel_err(o.x, s.x) < 1e-12 and CS.rel_err(o.lam, s.lam) < 1e-12
    for i, s in enumerate(singles):
        batch.set_dual_initialization(i, s.lam)
    assert batch.solve(qp_outs) == capi.TREEQP_OK
    assert all(o.iter == 0 for o in qp_outs)
    batch.destroy()


def test_sharded_solve_matches_single_gpu():
    """one tree QP sharded by subtree over 2 GPUs (NCCL all-gather of the cut nodes' Schur complements,
    all-reduce of the line-search scalars) equals the single-GPU solve; needs >= 2 GPUs"""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    world = 4 if n >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "run_sharded.py"), "9"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout and r.stdout.count(" OK") == 5, r.stdout[-3000:]


@pytest.mark.parametrize("dense", [False, True])
def test_device_kkt_matches_host_kkt(dense, product_lib):
    """treeqp_tdunes_kkt_resident (k_kkt) = tree_qp_out_max_KKT_res (tree_qp_common.c:769) of the same
    iterates: on a NON-converged state (2 iterations) the residual is O(1e-2) and must agree to
    rounding; on the converged state both are at the tolerance level"""
    L = product_lib
    if dense:
        prob, opts = CS.get_dense_case("random_vary_dims")
    else:
        prob, opts = CS.get_case("random_vary_dims")
    for max_iter in (2, 200):
        qp_in, qp_out = capi.QpIn(L, prob), capi.QpOut(L, prob)
        s = capi.Tdunes(L, qp_in, **dict(opts, maxIter=max_iter))
        s.set_dual_initialization(np.zeros(prob.n_lambda))
        st = s.solve(qp_out)
        host = float(L.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
        dev = s.kkt_resident()
        if max_iter == 2:
            assert st == capi.TREEQP_MAXIMUM_ITERATIONS_REACHED and host > 1e-6
            assert abs(dev - host) <= 1e-9 * host
        else:
            assert st == 0 and host < 1e-8 and dev < 1e-8
        s.destroy()


def test_device_kkt_batch(product_lib):
    L = product_lib
    probs = P.linear_chain_batch(16, seed=3)
    qp_ins = [capi.QpIn(L, p) for p in probs]
    qp_outs = [capi.QpOut(L, p) for p in probs]
    b = capi.TdunesBatch(L, qp_ins)
    assert b.solve(qp_outs) == capi.TREEQP_OK and (b.status == 0).all()
    dev = b.kkt_resident()
    host = np.array([L.lib.tree_qp_out_max_KKT_res(i.ptr, o.ptr) for i, o in zip(qp_ins, qp_outs)])
    assert dev.shape == (16,) and (dev < 1e-8).all() and (host < 1e-8).all()
    b.destroy()


@pytest.mark.parametrize("shape", [(3, 1, 4, 4, 1), (3, 1, 6, 8, 3), (1, 1, 7, 6, 4), (2, 2, 5, 12, 4), (2, 1, 3, 8, 3)])
def test_chain_kernels_odd_counts_match_oracle(shape, product_lib, monkeypatch):
    """chain segments with an ODD number of chains / chain parents (one half-warp of the paired stage
    kernel and of the two-chains-per-warp sweeps stays empty), a pure chain (md = 1 at the root), and
    nx = 12 (one chain per warp): same iterates as the oracle and as the per-depth kernels"""
    md, Nr, Nh, nx, nu = shape
    prob = P.random_tree(P.multistage_nk(md, Nr, Nh), nx, nu, seed=11 + Nh, ubound=0.3)
    opts = dict(maxIter=200)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert (sol.status, sol.iter) == (s["status"], s["iter"]) and sol.status == 0
    assert sol.stats["ls_iters"] == s["ls_iters"]
    check_against(sol, s, prob, tol=1e-9)
    monkeypatch.setenv("TREEQP_B200_NO_CHAIN", "1")
    monkeypatch.setenv("TREEQP_B200_NO_PAIR", "1")
    ref = capi.solve_tdunes(product_lib, prob, **opts)
    assert ref.iter == sol.iter
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert np.array_equal(getattr(sol, k), getattr(ref, k)), k + ": chain kernels are not bit-identical to the per-depth kernels"


def test_large_tree_matches_reference_golden(product_lib):
    """config 3's QP at depth 14 (32 767 nodes) against the unmodified reference: same status and Newton
    iteration count, identical active sets over ALL 65 532 input bounds, x/u/lam/mu to 1e-10 on every
    29th entry plus the exact sums (fixture made by tests/golden/make_golden.py large)"""
    prob, opts = CS.large_case()
    g = CS.large_golden()
    assert str(g["input_sha256"]) == CS.checksum(prob)
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    s = CS.large_summary(prob, dict(status=sol.status, iter=sol.iter, x=sol.x, u=sol.u, lam=sol.lam, mu_x=sol.mu_x, mu_u=sol.mu_u))
    assert (s["status"], s["iter"]) == (int(g["status"]), int(g["iter"]))
    for k in ("u_lower", "u_upper", "x_bound"):
        assert np.array_equal(s[k], g[k]), "active set differs: " + k
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k + "_sample"], g[k + "_sample"]) < TOL, k
        assert abs(s[k + "_sum"] - float(g[k + "_sum"])) <= 1e-9 * max(1.0, abs(float(g[k + "_sum"]))), k
    assert sol.kkt < 1e-8


def test_dma_from_registered_caller_buffers_matches_staged_path(product_lib):
    """Large single trees with caller buffers page-locked through treeqp_b200_host_register: A, B and the vectors of
    tree_qp_in travel as strided DMA copies in the caller's layout and are converted on the device (k_unpack_run), the
    results are assembled on the device in tree_qp_out's layout and written in place (k_gather_run).  Same bits as the
    staged path (host packer + pinned block), also after the caller refills the buffer."""
    prob = P.random_tree(P.binary_tree_nk(13), 12, 4, seed=8, ubound=0.3)
    ref = capi.solve_tdunes(product_lib, prob, maxIter=200)          # staged path
    assert ref.status == 0
    prob2 = P.random_tree(P.binary_tree_nk(13), 12, 4, seed=9, ubound=0.3)
    ref2 = capi.solve_tdunes(product_lib, prob2, maxIter=200)
    qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
    assert qp_in.pin() and qp_out.pin()
    solver = capi.Tdunes(product_lib, qp_in, maxIter=200)
    for rep in range(2):
        solver.set_dual_initialization(np.zeros(prob.n_lambda))
        assert solver.solve(qp_out) == 0
        assert qp_out.iter == ref.iter
        for k in ("x", "u", "lam", "mu_x", "mu_u"):
            assert np.array_equal(getattr(qp_out, k), getattr(ref, k)), (rep, k)
        # the DMA path moves the caller's rows (padding included), the staged path the packed arrays: the byte counts tell them apart
        assert solver.stats()["h2d_bytes"] != ref.stats["h2d_bytes"], "the DMA path did not run"
    qp_in.set_data(prob2)   # new data in the same (page-locked) buffer
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    assert solver.solve(qp_out) == 0
    assert qp_out.iter == ref2.iter
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert np.array_equal(getattr(qp_out, k), getattr(ref2, k)), ("refilled", k)
    solver.destroy()
    qp_in.mem.unpin(); qp_out.mem.unpin()
    # unregistered again: the staged path, same results
    solver = capi.Tdunes(product_lib, qp_in, maxIter=200)
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    assert solver.solve(qp_out) == 0 and np.array_equal(qp_out.x, ref2.x)
    assert solver.stats()["h2d_bytes"] == ref2.stats["h2d_bytes"]
    solver.destroy()


@pytest.mark.parametrize("reg", [dict(), dict(regType=1), dict(regType=2, regTol=0.5, regValue=1e-2)])
def test_reuse_of_unchanged_blocks_is_bit_identical(reg, product_lib, monkeypatch):
    """Wide levels re-use the factorisation of blocks whose subtree kept its active set (the GPU form of
    checkLastActiveSet, bf_block_reuse in tdunes_fused.cuh): only the right-hand side is redone.  Against
    TREEQP_B200_NO_SKIP=1 (every block re-factored in every iteration): the same bits everywhere, the same
    per-iteration line-search trial counts, the same count of regularised blocks (a large regTol makes the
    on-the-fly regularisation fire, so that re-used regularised factors are counted too)."""
    prob = P.random_tree(P.binary_tree_nk(13), 12, 4, seed=5, ubound=0.2)
    monkeypatch.setenv("TREEQP_B200_NO_SKIP", "1")
    a = capi.solve_tdunes(product_lib, prob, maxIter=200, **reg)
    monkeypatch.setenv("TREEQP_B200_NO_SKIP", "0")
    b = capi.solve_tdunes(product_lib, prob, maxIter=200, **reg)
    assert (a.status, a.iter) == (b.status, b.iter)
    assert a.stats["ls_iters"] == b.stats["ls_iters"]
    assert a.stats["n_reg"] == b.stats["n_reg"]
    if reg.get("regType") == 2: assert a.stats["n_reg"] > 0
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_affine_stage_update_matches_recomputation(product_lib, monkeypatch):
    """Large trees form the stage solution of most iterations from the affine model of the line search (k_step_affine:
    qmod0 += tau dq, no pass over A and B) and refresh it from lambda every 4th iteration.  Against recomputing it in
    every iteration (TREEQP_B200_AFFINE=0, what the reference does): same status, Newton iterations, per-iteration
    trial counts and active sets, solutions equal to rounding; and the reference golden of the depth-14 tree
    (test_large_tree_matches_reference_golden) runs on the affine path."""
    prob = P.random_tree(P.binary_tree_nk(13), 12, 4, seed=5, ubound=0.2)
    monkeypatch.setenv("TREEQP_B200_AFFINE", "0")
    a = capi.solve_tdunes(product_lib, prob, maxIter=200)
    monkeypatch.delenv("TREEQP_B200_AFFINE")
    b = capi.solve_tdunes(product_lib, prob, maxIter=200)
    assert (a.status, a.iter) == (b.status, b.iter) and a.status == 0
    assert a.stats["ls_iters"] == b.stats["ls_iters"]
    assert b.stats["kernel_launches"] < a.stats["kernel_launches"], "the affine step did not replace any stage pass"
    check_against(b, dict(x=a.x, u=a.u, lam=a.lam, mu_x=a.mu_x, mu_u=a.mu_u), prob, tol=1e-12)
    assert b.kkt < 1e-8


def test_batch_of_large_trees_matches_individual_solves(product_lib):
    """Two large trees as one forest: wide levels re-use unchanged blocks, the stage QPs come from the affine model
    between refreshes, and one tree finishes earlier than the other (its blocks / nodes drop out at their status check).
    Per tree exactly what treeqp_tdunes_solve returns for it alone."""
    probs = [P.random_tree(P.binary_tree_nk(12), 12, 4, seed=21, ubound=0.2), P.random_tree(P.binary_tree_nk(12), 12, 4, seed=22, ubound=0.35)]
    singles = [capi.solve_tdunes(product_lib, p, maxIter=200) for p in probs]
    assert singles[0].iter != singles[1].iter, "the two trees should need different iteration counts"
    qp_ins = [capi.QpIn(product_lib, p) for p in probs]
    qp_outs = [capi.QpOut(product_lib, p) for p in probs]
    batch = capi.TdunesBatch(product_lib, qp_ins, maxIter=200)
    assert batch.solve(qp_outs) == capi.TREEQP_OK
    for i, (s, o) in enumerate(zip(singles, qp_outs)):
        assert batch.status[i] == s.status == 0 and o.iter == s.iter, i
        for k in ("x", "u", "lam", "mu_x", "mu_u"):
            assert np.array_equal(getattr(o, k), getattr(s, k)), (i, k)
    batch.destroy()

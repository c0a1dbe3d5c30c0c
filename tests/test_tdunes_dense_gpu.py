"""GPU parity tests of the dense stage-QP path (the reference's qpOASES plug-in,
dual_Newton_tree_qpoases.c): treeqp_tdunes_solve of libtreeqp_b200.so with
qp_solver = TREEQP_QPOASES_SOLVER versus (1) goldens produced by the unmodified reference running
qpOASES_e, (2) the oracle port, (3) the compiled reference when oracle/_ref travelled to the box.

Bar: identical status, Newton iteration count, line-search trial counts, active sets; x, u, lambda,
mu within 1e-10 relative (FP64)."""
import numpy as np
import pytest

import cases as CS
import oracle_port as OP
from treeqp_b200 import capi, problems as P

pytestmark = pytest.mark.gpu
TOL = 1e-10


def check(sol, ref, prob, tol=TOL):
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(getattr(sol, k), ref[k]) < tol, k
    for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
        assert np.array_equal(np.abs(getattr(sol, k) - lo) < 1e-12, np.abs(ref[k] - lo) < 1e-12), "active set (lower): " + k
        assert np.array_equal(np.abs(getattr(sol, k) - hi) < 1e-12, np.abs(ref[k] - hi) < 1e-12), "active set (upper): " + k


@pytest.mark.parametrize("name", CS.DENSE_CASE_NAMES)
def test_dense_solve_matches_reference_golden(name, product_lib):
    prob, opts = CS.get_dense_case(name)
    g = CS.dense_golden(name)
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert sol.status == int(g["status"])
    assert sol.iter == int(g["iter"]), "Newton iteration count differs from the reference"
    check(sol, g, prob)
    assert sol.kkt < 1e-8
    assert sol.stats["kernel_launches"] > 0


@pytest.mark.parametrize("name", ["spring_mass_tight", "spring_mass_diag", "random_nx6_nu4_md3", "linear_chain_mixed"])
def test_dense_line_search_counts_match_oracle(name, product_lib):
    prob, opts = CS.get_dense_case(name)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert sol.iter == s["iter"]
    assert sol.stats["ls_iters"] == s["ls_iters"], "per-iteration line-search trial counts differ"
    assert sol.stats["n_reg"] == s["n_reg"]


def test_dense_equals_clipping_on_diagonal_weights(product_lib):
    """with diagonal Q, R and S = 0 the dense solver must reproduce the clipping path"""
    prob, _ = CS.get_case("spring_mass_tight")
    a = capi.solve_tdunes(product_lib, prob)
    b = capi.solve_tdunes(product_lib, prob, qp_solver=1)
    assert (a.status, a.iter) == (b.status, b.iter)
    check(b, dict(x=a.x, u=a.u, lam=a.lam, mu_x=a.mu_x, mu_u=a.mu_u), prob, tol=1e-9)


@pytest.mark.parametrize("opts", [dict(maxIter=7), dict(lineSearchBeta=0.5, lineSearchGamma=0.2), dict(regType=1), dict(termCondition=1)])
def test_dense_options_match_oracle(opts, product_lib):
    prob, base = CS.get_dense_case("spring_mass_tight")
    opts = dict(base, **opts)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert (sol.status, sol.iter) == (s["status"], s["iter"])
    # a run cut off by maxIter stops in the middle of an ill-conditioned path (regularised singular
    # blocks amplify rounding differences): mid-trajectory iterates agree to ~1e-5, converged ones to 1e-9
    check(sol, s, prob, tol=1e-9 if sol.status == 0 else 1e-4)


def test_dense_live_reference_when_available(product_lib, ref_lib):
    """config 5 shape (quadcopter: nx=6, nu=4, md=3) with a genuinely dense Hessian (S != 0)"""
    prob = P.densify(P.random_tree(P.multistage_nk(3, 2, 6), 6, 4, seed=21, ubound=0.4), 9, 0.1, 0.2)
    r = capi.solve_tdunes(ref_lib, prob, qp_solver=1, maxIter=200)
    s = capi.solve_tdunes(product_lib, prob, qp_solver=1, maxIter=200)
    assert (s.status, s.iter) == (r.status, r.iter)
    check(s, dict(x=r.x, u=r.u, lam=r.lam, mu_x=r.mu_x, mu_u=r.mu_u), prob)


def test_dense_batch(product_lib):
    """forest of independent dense QPs through treeqp_tdunes_solve_batch"""
    probs = [P.densify(P.random_tree(P.multistage_nk(2, 2, 5), 4, 2, seed=30 + i, ubound=0.3), 40 + i, 0.1, 0.1) for i in range(6)]
    singles = [capi.solve_tdunes(product_lib, p, qp_solver=1) for p in probs]
    qp_ins = [capi.QpIn(product_lib, p) for p in probs]
    qp_outs = [capi.QpOut(product_lib, p) for p in probs]
    batch = capi.TdunesBatch(product_lib, qp_ins, qp_solver=1)
    assert batch.solve(qp_outs) == capi.TREEQP_OK
    for s, o in zip(singles, qp_outs):
        assert o.iter == s.iter
        assert np.array_equal(o.x, s.x) and np.array_equal(o.lam, s.lam)
    batch.destroy()


@pytest.mark.parametrize("n", CS.REF_RANDOM_QP)
def test_reference_random_qp_goldens(n, product_lib):
    """the reference's OWN test vectors for this path (examples/random_qp_utils/data0N.c) with its own
    acceptance bar (examples/random_qp.c:196-254): |x - xopt|, |u - uopt| < 1e-12, KKT < 1e-12"""
    prob, xopt, uopt = CS.ref_random_qp(n)
    sol = capi.solve_tdunes(product_lib, prob, **CS.REF_RANDOM_QP_OPTS)
    assert sol.status == 0 and sol.iter <= 1
    assert sol.kkt < 1e-12
    assert float(np.max(np.abs(sol.x - xopt))) < 1e-12 and float(np.max(np.abs(sol.u - uopt))) < 1e-12


@pytest.mark.parametrize("name", ["quadcopter", "crane", "linear_chain"])
def test_chain_walk_variants_are_bit_identical(name, product_lib, monkeypatch):
    """config-5 trees on the runtime-dimension path: the prefetching chain walk (k_factor_chain_pf / k_backsolve_chain_pf),
    the compile-time-dimension level kernels (TREEQP_B200_NO_CHAIN_PF=1), the runtime-dimension level kernels
    (+ NO_CHAIN_DC=1), one launch per depth (+ NO_CHAIN=1) and the two-slot first line-search pass
    (TREEQP_B200_DENSE_FIRST_FULL=0) give the same bits, iteration and trial counts (all switches are read at create)"""
    if name == "quadcopter":
        prob = P.densify_scaled(P.quadcopter(3, 2, 20, i_mpc=14, x0=CS.QUAD_X0), 4)
    elif name == "crane":
        prob = P.densify_scaled(P.crane(3, 2, 20, x0=CS.CRANE_X0), 5)
    else:
        prob = P.linear_chain(8, 7, 3, 2, 12, seed=2)
        prob.umin[:] = -0.3; prob.umax[:] = 0.3; prob.xmin[:] = -1e12; prob.xmax[:] = 1e12
        prob = P.densify(prob, 6, 0.1, 0.0)
    opts = dict(qp_solver=1, maxIter=200)
    sol = capi.solve_tdunes(product_lib, prob, **opts)
    assert sol.status == 0 and sol.kkt < 1e-8
    variants = [dict(TREEQP_B200_NO_CHAIN_PF="1"), dict(TREEQP_B200_NO_CHAIN_PF="1", TREEQP_B200_NO_CHAIN_DC="1"),
                dict(TREEQP_B200_NO_CHAIN="1"), dict(TREEQP_B200_DENSE_FIRST_FULL="0")]
    for env in variants:
        for k, val in env.items(): monkeypatch.setenv(k, val)
        ref = capi.solve_tdunes(product_lib, prob, **opts)
        for k in env: monkeypatch.delenv(k)
        assert (ref.status, ref.iter) == (sol.status, sol.iter), env
        assert ref.stats["ls_iters"] == sol.stats["ls_iters"], env
        for k in ("x", "u", "lam", "mu_x", "mu_u"):
            assert np.array_equal(getattr(sol, k), getattr(ref, k)), "%s differs under %s" % (k, env)

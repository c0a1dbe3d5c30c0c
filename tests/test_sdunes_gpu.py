"""GPU parity tests: treeqp_sdunes_solve of libtreeqp_b200.so (one persistent CUDA kernel) through
the C ABI versus (1) goldens produced by the unmodified reference, (2) the oracle port on the same
inputs, (3) the compiled reference itself when oracle/_ref travelled to the box.

Bar (BASELINE.json north_star): identical status, Newton iteration count and active sets;
x, u, lambda, mu within 1e-10 relative (FP64)."""
import numpy as np
import pytest

import cases as CS
import oracle_port as OP
from treeqp_b200 import capi, problems as P

pytestmark = pytest.mark.gpu
TOL = 1e-10


def check_against(sol, ref, prob, tol=TOL, active_sets=True):
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(getattr(sol, k), ref[k]) < tol, k
    if active_sets:
        for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
            assert np.array_equal(getattr(sol, k) == lo, ref[k] == lo), "active set (lower) differs: " + k
            assert np.array_equal(getattr(sol, k) == hi, ref[k] == hi), "active set (upper) differs: " + k


@pytest.mark.parametrize("name", CS.SDUNES_CASE_NAMES)
def test_solve_matches_reference_golden(name, product_lib):
    prob, opts = CS.get_sdunes_case(name)
    g = CS.sdunes_golden(name)
    sol = capi.solve_sdunes(product_lib, prob, **opts)
    assert sol.status == int(g["status"])
    assert sol.iter == int(g["iter"]), "Newton iteration count differs from the reference"
    check_against(sol, g, prob, active_sets=int(g["status"]) == 0)
    if sol.status == 0:   # the reference's own checker, no worse than on the reference's solution
        assert sol.kkt < max(1e-8, 10.0 * float(g["kkt"]))
    assert sol.stats["kernel_launches"] == 1   # the whole Newton loop is one persistent kernel


@pytest.mark.parametrize("name", CS.SDUNES_CASE_NAMES)
def test_line_search_counts_and_scenario_duals_match_oracle(name, product_lib):
    prob, opts = CS.get_sdunes_case(name)
    s = OP.solve_sdunes_port(prob, **CS.port_opts(opts))
    qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
    solver = capi.Sdunes(product_lib, qp_in, **opts)
    solver.set_dual_initialization(np.zeros(solver.n_lambda), np.zeros(solver.n_mu))
    status = solver.solve(qp_out)
    st = solver.stats()
    lam_s, mu_s = solver.duals()
    solver.destroy()
    assert (status, qp_out.iter) == (s["status"], s["iter"])
    assert st["ls_iters"] == s["ls_iters"], "per-iteration line-search trial counts differ"
    assert st["n_reg"] == s["n_reg"], "number of regularised blocks differs"
    assert CS.rel_err(lam_s, s["lam_scen"]) < TOL and CS.rel_err(mu_s, s["mu_scen"]) < TOL


@pytest.mark.parametrize("opts", [dict(regType=1), dict(termCondition=0, stationarityTolerance=1e-14),
                                  dict(termCondition=1), dict(checkLastActiveSet=0),
                                  dict(lineSearchBeta=0.5, lineSearchGamma=0.2), dict(maxIter=2)])
def test_options_match_oracle(opts, product_lib):
    prob, _ = CS.get_sdunes_case("spring_mass_cfg2_u01")
    s = OP.solve_sdunes_port(prob, **CS.port_opts(opts))
    sol = capi.solve_sdunes(product_lib, prob, **opts)
    assert (sol.status, sol.iter) == (s["status"], s["iter"])
    check_against(sol, s, prob, tol=1e-9, active_sets=sol.status == 0)


@pytest.mark.parametrize("name", ["spring_mass_cfg2_u01", "linear_chain_u05", "spring_mass_u02_nocache"])
def test_dual_start_matches_live_reference(name, product_lib, ref_lib):
    """non-zero dual start through treeqp_sdunes_set_dual_initialization (scenarios.c:2094)"""
    prob, opts = CS.get_sdunes_case(name)
    Ns, Nh, Nr, md, nl = P.scenario_dims(prob)
    rng = np.random.default_rng(7)
    lam0, mu0 = rng.standard_normal(nl), rng.standard_normal(Ns * Nh * int(prob.nx[1]))
    r = capi.solve_sdunes(ref_lib, prob, lam0=lam0, mu0=mu0, **opts)
    s = capi.solve_sdunes(product_lib, prob, lam0=lam0, mu0=mu0, **opts)
    assert (s.status, s.iter) == (r.status, r.iter)
    check_against(s, dict(x=r.x, u=r.u, lam=r.lam, mu_x=r.mu_x, mu_u=r.mu_u), prob)


def test_workspace_fields_callers_read(product_lib):
    """Ns, Nh, Nr, md head the workspace (examples/solve_qp_json.cpp:165-188 reads them directly)"""
    prob, _ = CS.get_sdunes_case("spring_mass_cfg2")
    qp_in = capi.QpIn(product_lib, prob)
    solver = capi.Sdunes(product_lib, qp_in)
    assert (solver.Ns, solver.Nh, solver.Nr, solver.md) == (81, 20, 4, 3)
    assert solver.n_lambda == P.scenario_dims(prob)[4]
    solver.destroy()


def test_warm_start_and_repeatability(product_lib):
    """second solve from the optimal scenario duals stops at iteration 0; a cold re-solve is
    bit-identical (fixed-shape reductions, no atomics on the data path)"""
    prob, opts = CS.get_sdunes_case("linear_chain_u05")
    qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
    solver = capi.Sdunes(product_lib, qp_in, **opts)
    z = (np.zeros(solver.n_lambda), np.zeros(solver.n_mu))
    solver.set_dual_initialization(*z)
    assert solver.solve(qp_out) == 0
    x1, lam1, it1 = qp_out.x, qp_out.lam, qp_out.iter
    lam_s, mu_s = solver.duals()
    solver.set_dual_initialization(lam_s.copy(), mu_s.copy())
    assert solver.solve(qp_out) == 0
    assert qp_out.iter == 0 and it1 > 0
    assert CS.rel_err(qp_out.x, x1) < 1e-12
    solver.set_dual_initialization(*z)
    assert solver.solve(qp_out) == 0
    assert qp_out.iter == it1 and np.array_equal(qp_out.x, x1) and np.array_equal(qp_out.lam, lam1)
    solver.destroy()


@pytest.mark.parametrize("name", ["spring_mass_cfg2", "spring_mass_u02_nocache", "linear_chain_u05"])
def test_fixed_and_runtime_dimension_chains_identical(name, product_lib, monkeypatch):
    """the scenario chains instantiated for nx = 4 / 8 and the runtime-dimension instance
    (TREEQP_B200_GENERIC_NX=1, read at create) keep the same operation order: same iterates"""
    prob, opts = CS.get_sdunes_case(name)
    res = []
    for generic in ("0", "1"):
        monkeypatch.setenv("TREEQP_B200_GENERIC_NX", generic)
        qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
        solver = capi.Sdunes(product_lib, qp_in, **opts)
        solver.set_dual_initialization(np.zeros(solver.n_lambda), np.zeros(solver.n_mu))
        status = solver.solve(qp_out)
        res.append((status, qp_out.iter, solver.stats()["ls_iters"], qp_out.x.copy(), qp_out.u.copy(), qp_out.lam.copy()))
        solver.destroy()
    a, b = res
    assert a[:3] == b[:3]
    for p, q in zip(a[3:], b[3:]):
        assert CS.rel_err(p, q) < 1e-13


def test_sdunes_and_tdunes_agree(product_lib):
    """the two decompositions solve the same QP: primal solutions agree to solver tolerance"""
    prob, _ = CS.get_sdunes_case("spring_mass_cfg2")
    a = capi.solve_sdunes(product_lib, prob)
    b = capi.solve_tdunes(product_lib, prob)
    assert a.status == 0 and b.status == 0
    assert CS.rel_err(a.x, b.x) < 1e-7 and CS.rel_err(a.u, b.u) < 1e-7

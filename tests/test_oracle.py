"""CPU tests: the oracle port is pinned against the reference's goldens (and, when the compiled
reference is present, against the reference itself on the same inputs)."""
import numpy as np
import pytest

import cases as CS
import oracle_port as OP
from treeqp_b200 import capi


@pytest.mark.parametrize("name", CS.CASE_NAMES)
def test_port_matches_reference_golden(name):
    prob, opts = CS.get_case(name)
    g = CS.golden(name)
    assert str(g["input_sha256"]) == CS.checksum(prob), "generator drifted from the committed golden"
    s = OP.solve_port(prob, **CS.port_opts(opts))
    assert s["status"] == int(g["status"])
    assert s["iter"] == int(g["iter"])          # identical Newton iteration count
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k], g[k]) < 1e-10, k
    # identical active sets: a variable sits exactly on a bound in both or in neither
    for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
        assert np.array_equal(s[k] == lo, g[k] == lo) and np.array_equal(s[k] == hi, g[k] == hi)


@pytest.mark.parametrize("name", ["spring_mass_cfg1", "spring_mass_tight", "random_d6_u05", "random_vary_dims"])
def test_port_matches_live_reference(name, ref_lib):
    prob, opts = CS.get_case(name)
    r = capi.solve_tdunes(ref_lib, prob, **opts)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    assert (s["status"], s["iter"]) == (r.status, r.iter)
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k], getattr(r, k)) < 1e-10, k


def test_reference_golden_iteration_counts():
    """BASELINE.md parity targets: 3 (config 1), 58 (xmax[1]=0.2), 5 (config-2 tree)."""
    assert int(CS.golden("spring_mass_cfg1")["iter"]) == 3
    assert int(CS.golden("spring_mass_tight")["iter"]) == 58
    assert int(CS.golden("spring_mass_cfg2tree")["iter"]) == 5
    for n in CS.CASE_NAMES:
        assert float(CS.golden(n)["kkt"]) < 1e-8


@pytest.mark.parametrize("opts", [dict(regType=1), dict(termCondition=0, stationarityTolerance=1e-14),
                                  dict(termCondition=1), dict(lineSearchBeta=0.5, lineSearchGamma=0.2), dict(lineSearchBeta=0.8, lineSearchGamma=0.05)])
def test_port_options_vs_live_reference(opts, ref_lib):
    prob, _ = CS.get_case("spring_mass_tight")
    r = capi.solve_tdunes(ref_lib, prob, **opts)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    assert (s["status"], s["iter"]) == (r.status, r.iter)
    assert CS.rel_err(s["x"], r.x) < 1e-9 and CS.rel_err(s["lam"], r.lam) < 1e-9


def test_port_no_regularisation_on_regular_problem(ref_lib):
    """regType NONE is only well defined when no block is singular (with the x[1]<=0.2 variant the
    factorisation meets exactly-singular blocks and the outcome depends on rounding noise)."""
    prob, _ = CS.get_case("spring_mass_cfg1")
    r = capi.solve_tdunes(ref_lib, prob, regType=0)
    s = OP.solve_port(prob, regType=0)
    assert (s["status"], s["iter"]) == (r.status, r.iter) == (0, 3)
    assert CS.rel_err(s["x"], r.x) < 1e-10


# ------------------------------------------------------------------------------- SDUNES port
def _sd_tol(name):
    return 1e-10


@pytest.mark.parametrize("name", CS.SDUNES_CASE_NAMES)
def test_sdunes_port_matches_reference_golden(name):
    prob, opts = CS.get_sdunes_case(name)
    g = CS.sdunes_golden(name)
    assert str(g["input_sha256"]) == CS.checksum(prob), "generator drifted from the committed golden"
    s = OP.solve_sdunes_port(prob, **CS.port_opts(opts))
    assert s["status"] == int(g["status"])
    assert s["iter"] == int(g["iter"])
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k], g[k]) < _sd_tol(name), k
    if int(g["status"]) == 0:
        for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
            assert np.array_equal(s[k] == lo, g[k] == lo) and np.array_equal(s[k] == hi, g[k] == hi)


@pytest.mark.parametrize("name", ["spring_mass_cfg2_u01", "linear_chain_u05", "spring_mass_u02_nocache"])
def test_sdunes_port_matches_live_reference_with_dual_start(name, ref_lib):
    """non-zero dual start through treeqp_sdunes_set_dual_initialization (scenarios.c:2094)"""
    from treeqp_b200 import problems as P
    prob, opts = CS.get_sdunes_case(name)
    Ns, Nh, Nr, md, nl = P.scenario_dims(prob)
    rng = np.random.default_rng(7)
    lam0, mu0 = rng.standard_normal(nl), rng.standard_normal(Ns * Nh * int(prob.nx[1]))
    r = capi.solve_sdunes(ref_lib, prob, lam0=lam0, mu0=mu0, **opts)
    s = OP.solve_sdunes_port(prob, lam0=lam0, mu0=mu0, **CS.port_opts(opts))
    assert (s["status"], s["iter"]) == (r.status, r.iter)
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k], getattr(r, k)) < 1e-10, k


def test_sdunes_reference_golden_iteration_counts():
    """BASELINE.md parity targets for SDUNES: 3 iterations (config 1 data), 5 (config 2)."""
    assert int(CS.sdunes_golden("spring_mass_cfg1")["iter"]) == 3
    assert int(CS.sdunes_golden("spring_mass_cfg2")["iter"]) == 5


# ------------------------------------------------------------------ dense stage-QP path (F3)
@pytest.mark.parametrize("name", CS.DENSE_CASE_NAMES)
def test_dense_port_matches_reference_golden(name):
    """oracle/tdunes_port.c dense path (box QP restated in port_common.h) versus the unmodified
    reference running qpOASES_e: same status, Newton iterations, solution, multipliers, active sets"""
    prob, opts = CS.get_dense_case(name)
    g = CS.dense_golden(name)
    assert str(g["input_sha256"]) == CS.checksum(prob), "generator drifted from the committed golden"
    s = OP.solve_port(prob, **CS.port_opts(opts))
    assert s["status"] == int(g["status"])
    assert s["iter"] == int(g["iter"])
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k], g[k]) < 1e-10, k
    for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
        assert np.array_equal(np.abs(s[k] - lo) < 1e-12, np.abs(g[k] - lo) < 1e-12)
        assert np.array_equal(np.abs(s[k] - hi) < 1e-12, np.abs(g[k] - hi) < 1e-12)


@pytest.mark.parametrize("name", ["spring_mass_tight", "random_nx6_nu4_md3"])
def test_dense_port_matches_live_reference(name, ref_lib):
    prob, opts = CS.get_dense_case(name)
    opts = dict(opts, lineSearchBeta=0.5, lineSearchGamma=0.2)
    r = capi.solve_tdunes(ref_lib, prob, **opts)
    s = OP.solve_port(prob, **CS.port_opts(opts))
    assert (s["status"], s["iter"]) == (r.status, r.iter)
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        assert CS.rel_err(s[k], getattr(r, k)) < 1e-10, k


# --------------------------------------- the reference's own golden vectors (examples/random_qp.c)
def _max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) if len(a) else 0.0


@pytest.mark.parametrize("n", CS.REF_RANDOM_QP)
def test_port_reproduces_reference_random_qp_goldens(n):
    """examples/random_qp.c:196-254: max |x - xopt|, |u - uopt| < 1e-12, KKT < 1e-12, status 0, with
    the options of :131-140 (dense stage QPs, maxIter 10, tolerance 1e-10, no regularisation)"""
    prob, xopt, uopt = CS.ref_random_qp(n)
    s = OP.solve_port(prob, **CS.port_opts(CS.REF_RANDOM_QP_OPTS))
    assert s["status"] == 0 and s["iter"] <= 1
    assert _max_abs(s["x"], xopt) < 1e-12 and _max_abs(s["u"], uopt) < 1e-12


@pytest.mark.parametrize("n", CS.REF_RANDOM_QP)
def test_live_reference_reproduces_its_random_qp_goldens(n, ref_lib):
    """the compiled reference through our ctypes driver: proves the fixture + driver path"""
    prob, xopt, uopt = CS.ref_random_qp(n)
    r = capi.solve_tdunes(ref_lib, prob, **CS.REF_RANDOM_QP_OPTS)
    assert r.status == 0 and r.kkt < 1e-12
    assert _max_abs(r.x, xopt) < 1e-12 and _max_abs(r.u, uopt) < 1e-12

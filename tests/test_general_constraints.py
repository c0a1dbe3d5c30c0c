"""General stage constraints dmin <= C x + D u <= dmax (the reference's QProblem path,
dual_Newton_tree_qpoases.c:181-193, 288-298, 541-565).

CPU: the containers (setters, tree_qp_out.mu_d, the general-constraint terms of tree_qp_out_max_KKT_res)
against the compiled reference.  GPU: treeqp_tdunes_solve with nc > 0 against goldens made by the
unmodified reference (tests/golden/make_golden.py general) -- same status, Newton iteration count, active
rows, x/u/lam/mu to 1e-10 -- and, where the reference's own nc > 0 path does not converge (it does not on
every QP, see tests/cases.py::general_cases), through properties the QP itself offers: the KKT residual of
tree_qp_out_max_KKT_res, and equality with the bounded twin when the rows only restate bounds."""
import numpy as np
import pytest

import cases as CS
from treeqp_b200 import capi, problems as P

TOL = 1e-10


def _solve(lib, prob, **opts):
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    s = capi.Tdunes(lib, qp_in, **opts)
    s.set_dual_initialization(np.zeros(prob.n_lambda))
    st = s.solve(qp_out)
    kkt = float(lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
    out = dict(status=st, iter=qp_out.iter, x=qp_out.x, u=qp_out.u, lam=qp_out.lam, mu_x=qp_out.mu_x, mu_u=qp_out.mu_u,
               mu_d=qp_out.mu_d, kkt=kkt, qp_in=qp_in, qp_out=qp_out)
    if lib.is_product:
        s.destroy()
    return out


def test_kkt_function_with_general_constraints_matches_reference(product_lib, ref_lib):
    """our tree_qp_out_max_KKT_res on the reference's solution (containers are layout-compatible), with
    the multipliers perturbed so that every general-constraint term is exercised"""
    prob, opts = CS.get_general_case("spring_mass_dense_rows")
    r = _solve(ref_lib, prob, **opts)
    assert r["status"] == 0 and r["kkt"] < 1e-8
    ours = float(product_lib.lib.tree_qp_out_max_KKT_res(r["qp_in"].ptr, r["qp_out"].ptr))
    assert abs(ours - r["kkt"]) <= 1e-12 + 1e-9 * r["kkt"]
    # the same QP built by OUR containers, evaluated by the reference's function on the same solution vectors
    g = CS.general_golden("spring_mass_dense_rows")
    assert str(g["input_sha256"]) == CS.checksum(prob)
    qi = capi.QpIn(product_lib, prob)
    assert float(ref_lib.lib.tree_qp_out_max_KKT_res(qi.ptr, r["qp_out"].ptr)) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("name", CS.GENERAL_CASE_NAMES)
def test_general_constraints_match_reference_golden(name, product_lib):
    prob, opts = CS.get_general_case(name)
    g = CS.general_golden(name)
    assert str(g["input_sha256"]) == CS.checksum(prob)
    s = _solve(product_lib, prob, **opts)
    assert s["status"] == int(g["status"]) and s["iter"] == int(g["iter"]), "status / Newton iterations differ from the reference"
    for k in ("x", "u", "lam", "mu_x", "mu_u", "mu_d"):
        assert CS.rel_err(s[k], g[k]) < TOL, k
    assert np.array_equal(s["mu_d"] > 0, g["mu_d"] > 0) and np.array_equal(s["mu_d"] < 0, g["mu_d"] < 0), "active rows differ"
    assert s["kkt"] < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["crane_cfg5", "random_nx6_nu4_md3"])
def test_unit_rows_equal_the_bounded_twin(name, product_lib):
    """conversions on which the REFERENCE's nc > 0 path fails to converge (u[0] only): the QP is the same as its
    bounded twin, so the solution and the Newton iteration count must be the twin's.  (quadcopter_cfg5 with this
    conversion runs into maxIter on the reference AND here: weights spanning 5e5 make the Armijo test sensitive to
    the last bits of the projector, which the Schur-complement form P_box - W S^-1 W' does not reproduce.)"""
    prob, opts = CS.get_dense_case(name)
    twin = _solve(product_lib, prob, **opts)
    rows = _solve(product_lib, P.bounds_to_general(prob, (), (0,)), **opts)
    assert (rows["status"], rows["iter"]) == (twin["status"], twin["iter"]) and twin["status"] == 0
    for k in ("x", "u", "lam"):
        assert CS.rel_err(rows[k], twin[k]) < 1e-9, k
    assert rows["kkt"] < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("width", [4.0, 1.0, 0.7])
def test_coupled_rows_satisfy_kkt(width, product_lib):
    """rows that couple states and inputs (c, d ~ U(-1, 1), two-sided): no reference answer exists (its nc > 0
    path runs into maxIter on these), the certificate is the KKT residual of tree_qp_out_max_KKT_res"""
    base = P.densify(P.random_tree(P.multistage_nk(3, 2, 6), 5, 3, seed=6, ubound=0.5), 2, 0.1, 0.1)
    prob = P.add_general_rows(base, 3, 2, width)
    s = _solve(product_lib, prob, qp_solver=1, maxIter=200)
    assert s["status"] == 0 and s["kkt"] < 1e-8, (s["status"], s["iter"], s["kkt"])
    if width >= 4.0:   # wide rows stay inactive: the solution is the one without rows
        free = _solve(product_lib, base, qp_solver=1, maxIter=200)
        assert s["iter"] == free["iter"] and CS.rel_err(s["x"], free["x"]) < 1e-10 and not np.any(s["mu_d"])
    else:
        assert np.any(s["mu_d"] != 0), "the narrow rows must be active somewhere"

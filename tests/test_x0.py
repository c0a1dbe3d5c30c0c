"""Initial-state handling (the MPC re-solve path).

CPU: the host copies of tree_qp_in_eliminate_x0 / tree_qp_in_set_x0_colmaj (tree_qp_common.c:404-525,
2154-2235) against the reference's own, by solving what each container built with the REFERENCE
solver.  GPU: treeqp_tdunes_update_x0 + warm_start_resident against full uploads."""
import numpy as np
import pytest

import cases as CS
from treeqp_b200 import capi, problems as P

X0A = [0.01, 0.02, 0.03, 0.04]
X0B = [0.3, 0.5, -0.4, 0.2]


def _ref_solve(ref, qp_in, prob_e):
    qp_out = capi.QpOut(ref, prob_e)
    s = capi.Tdunes(ref, qp_in)
    s.set_dual_initialization(np.zeros(prob_e.n_lambda))
    st = s.solve(qp_out)
    return st, qp_out.iter, qp_out.x, qp_out.u, qp_out.lam


def test_eliminate_and_set_x0_match_reference(product_lib, ref_lib):
    """same QP, x0 eliminated and then changed, once through our containers and once through the
    reference's; both solved by the reference's TDUNES"""
    prob = P.spring_mass(3, 2, 10, x0=X0A)
    prob_e = P.eliminate_x0(prob)
    res = {}
    for name, lib in (("ours", product_lib), ("ref", ref_lib)):
        qi = capi.QpIn(lib, prob)
        qi.eliminate_x0()
        first = _ref_solve(ref_lib, qi, prob_e)
        qi.set_x0(X0B)
        second = _ref_solve(ref_lib, qi, prob_e)
        res[name] = (first, second)
    for a, b in zip(res["ours"], res["ref"]):
        assert a[0] == b[0] == 0 and a[1] == b[1]
        for va, vb in zip(a[2:], b[2:]):
            assert CS.rel_err(va, vb) < 1e-12
    # and against the numpy elimination used by the other tests
    qi = capi.QpIn(ref_lib, P.eliminate_x0(P.spring_mass(3, 2, 10, x0=X0B)))
    chk = _ref_solve(ref_lib, qi, prob_e)
    assert chk[1] == res["ours"][1][1]
    for va, vb in zip(chk[2:], res["ours"][1][2:]):
        assert CS.rel_err(va, vb) < 1e-10


def test_set_x0_on_pinned_root(product_lib):
    """nx[0] > 0: set_x0 writes the root's bounds (tree_qp_common.c:2161-2171)"""
    prob = P.spring_mass(3, 2, 10, x0=X0A)
    qi = capi.QpIn(product_lib, prob)
    qi.set_x0(X0B)
    # solving needs a GPU; here only the container: rebuild a Problem view through a second set_x0
    qi.set_x0(X0A)  # idempotent path, no crash, dimensions respected


@pytest.mark.gpu
@pytest.mark.parametrize("eliminated", [True, False])
def test_update_x0_matches_full_upload(eliminated, product_lib):
    """closed-loop pattern: upload once, then per sample set_x0 + update_x0 + solve_resident"""
    L = product_lib
    prob = P.spring_mass(3, 2, 10, x0=X0A)
    dims = P.eliminate_x0(prob) if eliminated else prob
    qi = capi.QpIn(L, prob)
    if eliminated:
        qi.eliminate_x0()
    qo = capi.QpOut(L, dims)
    s = capi.Tdunes(L, qi)
    lam0 = np.zeros(dims.n_lambda)
    s.set_dual_initialization(lam0)
    assert s.upload() == capi.TREEQP_OK
    full_bytes = s.stats()["h2d_bytes"]
    assert s.solve_resident() == 0
    # new initial state: only the root and its children travel
    qi.set_x0(X0B)
    assert s.update_x0() == capi.TREEQP_OK
    assert s.stats()["h2d_bytes"] < full_bytes / 5
    s.set_dual_initialization(lam0)
    assert s.solve_resident() == 0
    assert s.download(qo) == capi.TREEQP_OK
    x, u, lam, it = qo.x, qo.u, qo.lam, s.stats()["iter"]
    # the same QP built from scratch with the new x0
    pb = P.spring_mass(3, 2, 10, x0=X0B)
    ref = capi.solve_tdunes(L, P.eliminate_x0(pb) if eliminated else pb)
    assert ref.status == 0 and ref.iter == it
    for a, b in ((x, ref.x), (u, ref.u), (lam, ref.lam)):
        assert CS.rel_err(a, b) < 1e-10
    # warm start from the multipliers left in HBM: nothing to do
    assert s.warm_start_resident() == capi.TREEQP_OK
    assert s.solve_resident() == 0
    assert s.stats()["iter"] == 0
    assert s.download(qo) == capi.TREEQP_OK
    assert CS.rel_err(qo.x, x) < 1e-12
    s.destroy()


@pytest.mark.gpu
def test_warm_start_survives_update_x0(product_lib):
    """warm_start_resident BEFORE update_x0: the x0 update must not overwrite the device-resident dual
    start of the root's children with the stale host copy (zeros)"""
    L = product_lib
    prob = P.spring_mass(3, 2, 10, x0=X0A, xmax1=0.2)
    qi = capi.QpIn(L, prob)
    qi.eliminate_x0()
    dims = P.eliminate_x0(prob)
    qo = capi.QpOut(L, dims)
    s = capi.Tdunes(L, qi)
    s.set_dual_initialization(np.zeros(dims.n_lambda))
    assert s.upload() == capi.TREEQP_OK and s.solve_resident() == 0
    cold_iters = s.stats()["iter"]
    assert s.download(qo) == capi.TREEQP_OK
    lam, x = qo.lam.copy(), qo.x.copy()
    # order A (the one that used to lose the warm start): warm start, THEN the x0 update (same x0)
    assert s.warm_start_resident() == capi.TREEQP_OK
    qi.set_x0(X0A)
    assert s.update_x0() == capi.TREEQP_OK
    assert s.solve_resident() == 0
    assert s.stats()["iter"] == 0, "dual start was overwritten by update_x0 (cold start needs %d iterations)" % cold_iters
    assert s.download(qo) == capi.TREEQP_OK
    assert CS.rel_err(qo.lam, lam) < 1e-14 and CS.rel_err(qo.x, x) < 1e-12
    s.destroy()


@pytest.mark.gpu
def test_second_solve_continues_from_last_multipliers(product_lib, ref_lib):
    """The reference iterates work->slambda in place (dual_Newton_tree.c:966): treeqp_tdunes_solve called
    again WITHOUT set_dual_initialization starts from the multipliers of the previous solve.  Same
    sequence of calls on both libraries, truncated first solve so that the second one has work to do."""
    prob = P.spring_mass(3, 2, 10, x0=X0A, xmax1=0.2)
    res = {}
    for name, lib in (("ours", product_lib), ("ref", ref_lib)):
        qi, qo = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
        s = capi.Tdunes(lib, qi, maxIter=10)
        s.set_dual_initialization(np.zeros(prob.n_lambda))
        seq = []
        for _ in range(7):   # 58 Newton iterations in slices of 10: (1, 10) x 5, (0, 8), (0, 0)
            st = s.solve(qo)
            seq.append((int(st), int(qo.iter)))
        res[name] = (seq, qo.x.copy(), qo.lam.copy())
        if lib is product_lib:
            s.destroy()
    assert res["ours"][0] == res["ref"][0], (res["ours"][0], res["ref"][0])
    assert res["ref"][0][0] == (1, 10) and res["ref"][0][-1] == (0, 0)
    assert CS.rel_err(res["ours"][1], res["ref"][1]) < 1e-10 and CS.rel_err(res["ours"][2], res["ref"][2]) < 1e-10


@pytest.mark.gpu
def test_batch_update_x0_matches_fresh_batch(product_lib):
    """closed-loop step of a batch: set_x0 on every QP + treeqp_tdunes_batch_update_x0 (one strided copy per array)
    against a batch built from scratch with the new initial states"""
    L = product_lib
    n = 12
    rng = np.random.default_rng(3)
    def make(x0s):
        probs = [P.spring_mass(3, 2, 10, x0=x0) for x0 in x0s]
        qis = [capi.QpIn(L, p) for p in probs]
        for q in qis:
            q.eliminate_x0()
        dims = [P.eliminate_x0(p) for p in probs]
        return probs, qis, dims
    x0a = [np.array(X0A) + 0.01 * rng.standard_normal(4) for _ in range(n)]
    x0b = [np.array(X0B) * (0.5 + 0.5 * rng.random()) for _ in range(n)]
    _, qis, dims = make(x0a)
    qos = [capi.QpOut(L, d) for d in dims]
    b = capi.TdunesBatch(L, qis)
    assert b.upload() == capi.TREEQP_OK and b.solve_resident() == capi.TREEQP_OK
    full = b.stats()["h2d_bytes"]
    for q, x0 in zip(qis, x0b):
        q.set_x0(x0)
    assert b.update_x0() == capi.TREEQP_OK
    assert 0 < b.stats()["h2d_bytes"] < full / 4
    assert b.solve_resident() == capi.TREEQP_OK and b.download(qos) == capi.TREEQP_OK
    it_upd = [o.iter for o in qos]
    x_upd = [o.x.copy() for o in qos]
    _, qis2, dims2 = make(x0b)
    qos2 = [capi.QpOut(L, d) for d in dims2]
    b2 = capi.TdunesBatch(L, qis2)
    assert b2.solve(qos2) == capi.TREEQP_OK
    assert (b.status == 0).all() and (b2.status == 0).all()
    for i in range(n):
        assert qos2[i].iter == it_upd[i]
        assert CS.rel_err(x_upd[i], qos2[i].x) < 1e-12
    b.destroy(); b2.destroy()

"""Seeded parity cases shared by the CPU (oracle) and GPU tests; goldens in tests/golden/."""
import hashlib
import os

import numpy as np

from treeqp_b200 import problems as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# closed-loop states a few samples into the manoeuvre: input (quadcopter) / state (crane) bounds active,
# and the reference meets its own KKT bar with both stage solvers (at x0 = 0 its qpOASES path stops after
# one iteration with a KKT residual of 0.46 on the crane, so x0 = 0 cannot serve as a golden)
QUAD_X0 = [0.03, -0.03, 0.01, 0.8, -0.8, 0.3]
CRANE_X0 = [0.0, 0.2, 0.1, -0.3]


def cases():
    yield "spring_mass_cfg1", P.spring_mass(3, 2, 10), {}
    yield "spring_mass_tight", P.spring_mass(3, 2, 10, xmax1=0.2), {}
    yield "spring_mass_cfg2tree", P.spring_mass(3, 4, 20), {}
    yield "random_d6_u05", P.random_tree(P.binary_tree_nk(6), 12, 4, seed=1, ubound=0.5), dict(maxIter=200)
    yield "random_d10_u02", P.random_tree(P.binary_tree_nk(10), 12, 4, seed=2, ubound=0.2), dict(maxIter=200)
    yield "random_d8_u05", P.random_tree(P.binary_tree_nk(8), 12, 4, seed=3, ubound=0.5), dict(maxIter=200)
    yield "random_vary_dims", P.random_tree(P.multistage_nk(3, 2, 5), 5, 3, seed=4, ubound=0.3, vary_dims=True), dict(maxIter=200)
    yield "random_ternary_chain", P.random_tree(P.multistage_nk(3, 2, 8), 6, 2, seed=5, ubound=0.2, xbound=3.0), dict(maxIter=200)
    yield "linear_chain_cfg4", P.linear_chain(4, 3, 3, 2, 20, seed=0), {}
    yield "linear_chain_cfg4_s7", P.linear_chain(4, 3, 3, 2, 20, seed=7), {}
    yield "linear_chain_pinned_x0", P.linear_chain(2, 1, 2, 2, 6, seed=1, eliminate_x0=False), {}
    # config 5's benchmark models (benchmark/quadcopter, benchmark/crane), diagonal weights, clipping
    yield "quadcopter_cfg5", P.quadcopter(3, 2, 20, i_mpc=14, x0=QUAD_X0), dict(maxIter=200)
    yield "crane_cfg5", P.crane(3, 2, 20, x0=CRANE_X0), dict(maxIter=200)


def dense_cases():
    """Dense stage-QP path (qp_solver = TREEQP_QPOASES_SOLVER, dual_Newton_tree_qpoases.c): name,
    problem, options (qp_solver is per node or a scalar)."""
    D = P.densify
    Q = dict(qp_solver=1)
    yield "spring_mass_cfg1", D(P.spring_mass(3, 2, 10), 1, 0.05, 0.0), dict(Q)
    yield "spring_mass_tight", D(P.spring_mass(3, 2, 10, xmax1=0.2), 1, 0.05, 0.05), dict(Q)
    yield "spring_mass_diag", P.spring_mass(3, 2, 10, xmax1=0.2), dict(Q)      # diagonal weights, dense solver
    yield "random_unconstrained", D(P.random_tree(P.multistage_nk(2, 2, 4), 5, 3, seed=6, ubound=1e12, vary_dims=True), 2, 0.2, 0.3), dict(Q)
    yield "random_vary_dims", D(P.random_tree(P.binary_tree_nk(5), 5, 2, seed=7, ubound=0.5, vary_dims=True), 3, 0.1, 0.1), dict(Q)
    yield "random_nx6_nu4_md3", D(P.random_tree(P.multistage_nk(3, 2, 8), 6, 4, seed=8, ubound=0.3), 4, 0.1, 0.0), dict(Q, maxIter=200)
    yield "random_nx4_nu1_xbox", D(P.random_tree(P.multistage_nk(3, 2, 8), 4, 1, seed=9, ubound=0.2, xbound=3.0), 5, 0.05, 0.0), dict(Q, maxIter=200)
    yield "linear_chain_nm8_md4", D(_ubox(P.linear_chain(8, 7, 4, 2, 5, seed=2), 0.3, 1e12), 6, 0.1, 0.0), dict(Q)
    # config 5 as named: the quadcopter and crane trees with dense stage Hessians (qpOASES stage solver)
    yield "quadcopter_cfg5", P.densify_scaled(P.quadcopter(3, 2, 20, i_mpc=14, x0=QUAD_X0), 4), dict(Q, maxIter=200)
    yield "crane_cfg5", P.densify_scaled(P.crane(3, 2, 20, x0=CRANE_X0), 5), dict(Q, maxIter=200)
    yield "quadcopter_cfg5_diag", P.quadcopter(3, 2, 20, i_mpc=12), dict(Q, maxIter=200)
    lc = _ubox(P.linear_chain(4, 3, 3, 2, 8, seed=3), 0.4, 1e12)
    # per-node choice: dense solver on the parents, clipping on the leaves (the reference's own
    # workspace carving only supports mixes whose dense nodes come first: create_strmat asserts
    # 64-byte alignment after a clipping node's odd-sized data, treeqp/utils/memory.c:117)
    yield "linear_chain_mixed", lc, dict(qp_solver=[1 if lc.nk[k] > 0 else 0 for k in range(lc.Nn)])


def general_cases():
    """General stage constraints dmin <= C x + D u <= dmax on dense-solver nodes (the reference's
    QProblem path, dual_Newton_tree_qpoases.c:181-193, 288-298).  The reference's own data for this path
    (examples/spring_mass.c:100-223) turns bounds into unit rows; the same is done here on the dense cases on
    which the reference's nc > 0 path converges to the solution of the bounded QP (it does not on every QP:
    on e.g. quadcopter_cfg5 with only u[0] converted it runs into maxIter with a KKT residual of 0.3 while the
    bounded twin takes 17 iterations -- those conversions cannot be goldens)."""
    G = P.bounds_to_general
    Q = dict(qp_solver=1, maxIter=200)
    yield "spring_mass_rows", G(P.spring_mass(3, 2, 10, xmax1=0.2), (1,), (0,)), dict(Q)
    yield "spring_mass_dense_rows", G(P.densify(P.spring_mass(3, 2, 10, xmax1=0.2), 1, 0.05, 0.05), (1,), (0,)), dict(Q)
    for name, xi, ui in (("random_nx6_nu4_md3", (), (0, 1)), ("random_nx4_nu1_xbox", (1,), (0,)), ("linear_chain_nm8_md4", (), (0, 1)),
                         ("quadcopter_cfg5", (), (0, 1)), ("crane_cfg5", (1,), (0,)), ("random_vary_dims", (1,), (0,))):
        prob, opts = get_dense_case(name)
        yield name + "_rows", G(prob, xi, ui), dict(opts)


GENERAL_CASE_NAMES = ["spring_mass_rows", "spring_mass_dense_rows", "random_nx6_nu4_md3_rows", "random_nx4_nu1_xbox_rows",
                      "linear_chain_nm8_md4_rows", "quadcopter_cfg5_rows", "crane_cfg5_rows", "random_vary_dims_rows"]


def get_general_case(name):
    for n, p, o in general_cases():
        if n == name:
            return p, o
    raise KeyError(name)


def general_golden(name):
    return np.load(os.path.join(GOLDEN, "tdunes_general_%s.npz" % name))


def get_dense_case(name):
    for n, p, o in dense_cases():
        if n == name:
            return p, o
    raise KeyError(name)


def dense_golden(name):
    return np.load(os.path.join(GOLDEN, "tdunes_dense_%s.npz" % name))


def _ubox(prob, ub, xb=None):
    prob.umin[:] = -ub
    prob.umax[:] = ub
    if xb is not None:
        prob.xmin[:] = -xb
        prob.xmax[:] = xb
    return prob


_X0B = [0.3, 0.5, -0.4, 0.2]

# the reference's OWN golden vectors for the dense TDUNES path (examples/random_qp_utils/data0N.c,
# checked by examples/random_qp.c:196-254); fixtures made by tests/golden/make_random_qp_fixtures.py
REF_RANDOM_QP = ["%02d" % n for n in range(6)]
REF_RANDOM_QP_OPTS = dict(qp_solver=1, maxIter=10, stationarityTolerance=1e-10, regType=0)  # random_qp.c:131-140


def ref_random_qp(n):
    """(problem, xopt, uopt) of examples/random_qp_utils/data<n>.c: dense Q/R/S, infinite bounds
    (tree_qp_in_set_inf_bounds, +-TREEQP_INF = 1e12), irregular nx/nu"""
    from treeqp_b200.capi import Problem
    d = np.load(os.path.join(GOLDEN, "refdata_random_qp_%s.npz" % n))
    X, U = int(d["nx"].sum()), int(d["nu"].sum())
    prob = Problem(nk=d["nk"].astype(np.int32), nx=d["nx"].astype(np.int32), nu=d["nu"].astype(np.int32),
                   A=d["A"], B=d["B"], b=d["b"], q=d["q"], r=d["r"], Q=d["Q"], R=d["R"], S=d["S"],
                   xmin=np.full(X, -1e12), xmax=np.full(X, 1e12), umin=np.full(U, -1e12), umax=np.full(U, 1e12),
                   meta=dict(source="examples/random_qp_utils/data%s.c" % n))
    return prob, d["xopt"], d["uopt"]


def sdunes_cases():
    """SDUNES parity cases (x0 eliminated).  The reference's SDUNES does not converge once state
    bounds make the scenario Hessians singular (measured: 100 iterations on every such probe), so
    the converging cases bound the inputs.  (On those degenerate paths the reference produces
    directions of magnitude 1e17 out of regularised singular systems: the outcome is rounding
    noise, there is nothing to pin.)"""
    E = P.eliminate_x0
    yield "spring_mass_cfg1", E(P.spring_mass(3, 2, 10)), {}
    yield "spring_mass_cfg2", E(P.spring_mass(3, 4, 20)), {}
    yield "spring_mass_cfg2_u01", _ubox(E(P.spring_mass(3, 4, 20, x0=_X0B)), 0.1, 1e12), {}
    yield "spring_mass_u02_nocache", _ubox(E(P.spring_mass(3, 2, 10, x0=_X0B)), 0.2, 1e12), dict(checkLastActiveSet=0)
    yield "linear_chain_u05", _ubox(P.linear_chain(4, 3, 3, 2, 20, seed=0), 0.5, 1e12), {}
    yield "linear_chain_nm8_md4", _ubox(P.linear_chain(8, 7, 4, 2, 20, seed=2), 0.3, 1e12), {}
    yield "linear_chain_md2_Nr3", _ubox(P.linear_chain(2, 1, 2, 3, 10, seed=1), 0.3, 1e12), {}
    yield "linear_chain_md3_Nr3", _ubox(P.linear_chain(3, 2, 3, 3, 12, seed=2), 0.3, 1e12), dict(regType=1)
    # gamma > 0.5 rejects full Newton steps even on a quadratic: exercises the backtracking.  The
    # tolerance is 1e-6 because below ~1e-8 the Armijo test compares differences of 1e-16 (noise).
    yield "linear_chain_u05_ls", _ubox(P.linear_chain(4, 3, 3, 2, 20, seed=4), 0.5, 1e12), dict(lineSearchGamma=0.6, lineSearchBeta=0.7, stationarityTolerance=1e-6)
    yield "spring_mass_cfg1_ls", E(P.spring_mass(3, 2, 10)), dict(lineSearchGamma=0.55, lineSearchBeta=0.8, stationarityTolerance=1e-6)


CASE_NAMES = [c[0] for c in cases()]
DENSE_CASE_NAMES = [c[0] for c in dense_cases()]
SDUNES_CASE_NAMES = [c[0] for c in sdunes_cases()]


def get_sdunes_case(name):
    for n, p, o in sdunes_cases():
        if n == name:
            return p, o
    raise KeyError(name)


def sdunes_golden(name):
    return np.load(os.path.join(GOLDEN, "sdunes_%s.npz" % name))


def get_case(name):
    for n, p, o in cases():
        if n == name:
            return p, o
    raise KeyError(name)


def checksum(prob):
    h = hashlib.sha256()
    for name in ["nk", "nx", "nu", "A", "B", "b", "Qd", "Rd", "Q", "R", "S", "q", "r", "xmin", "xmax", "umin", "umax",
                 "nc", "C", "D", "dmin", "dmax"]:
        if getattr(prob, name) is not None:
            h.update(np.ascontiguousarray(getattr(prob, name)).tobytes())
    return h.hexdigest()


def golden(name):
    return np.load(os.path.join(GOLDEN, "tdunes_%s.npz" % name))


# ---- large-tree golden: config 3's QP two depths below the benchmark size (32 767 nodes, 21 Newton
# iterations, 6 s on the reference).  The solution vectors are 12 MB, so the fixture keeps every
# LARGE_STRIDE-th entry of each vector, the packed active-set masks of ALL bounds and the exact sums.
LARGE_NAME, LARGE_DEPTH, LARGE_STRIDE = "random_d14_u02", 14, 29


def large_case():
    return P.random_tree(P.binary_tree_nk(LARGE_DEPTH), 12, 4, seed=0, ubound=0.2), dict(maxIter=200)


def large_summary(prob, sol):
    """what the fixture stores of a solution (dict with status, iter, x, u, lam, mu_x, mu_u)"""
    out = {"status": int(sol["status"]), "iter": int(sol["iter"])}
    for k in ("x", "u", "lam", "mu_x", "mu_u"):
        v = np.asarray(sol[k])
        out[k + "_sample"] = v[::LARGE_STRIDE].copy()
        out[k + "_sum"] = float(np.sum(v))
        out[k + "_absmax"] = float(np.max(np.abs(v)))
    out["u_lower"] = np.packbits(np.asarray(sol["u"]) == prob.umin)
    out["u_upper"] = np.packbits(np.asarray(sol["u"]) == prob.umax)
    out["x_bound"] = np.packbits((np.asarray(sol["x"]) == prob.xmin) | (np.asarray(sol["x"]) == prob.xmax))
    return out


def large_golden():
    return np.load(os.path.join(GOLDEN, "tdunes_large_%s.npz" % LARGE_NAME))


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


PORT_OPT_NAMES = {"qp_solver", "checkLastActiveSet", "maxIter", "termCondition", "regType", "regTol", "regValue", "lineSearchMaxIter",
                  "lineSearchGamma", "lineSearchBeta"}


def port_opts(opts):
    o = {k: v for k, v in opts.items() if k in PORT_OPT_NAMES}
    if "stationarityTolerance" in opts:
        o["tol"] = opts["stationarityTolerance"]
    return o

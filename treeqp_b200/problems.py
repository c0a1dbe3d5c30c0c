"""Synthetic tree-QP generators for the parity tests and the benchmark.

Harness code: produces `capi.Problem` objects (flat column-major arrays) that are then pushed
through the C API's setters.  Every generator is seeded (numpy PCG64), so the reference, the
oracle port and the GPU path see byte-identical inputs.
"""
from __future__ import annotations

import os

import numpy as np

from .capi import Problem, TREEQP_INF

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def multistage_nk(md: int, Nr: int, Nh: int) -> np.ndarray:
    """children counts of a tree branching md-fold for Nr stages, then Nh-Nr chain stages
    (treeqp/utils/tree.c:247-280)."""
    nk, width = [], 1
    for stage in range(Nh):
        fan = md if stage < Nr else 1
        nk += [fan] * width
        width *= fan
    nk += [0] * width
    return np.asarray(nk, dtype=np.int32)


def tree_tables(nk: np.ndarray):
    """dad, stage, realisation id (treeqp/utils/tree.c:171-243) for breadth-first children counts."""
    Nn = len(nk)
    dad = np.full(Nn, -1, dtype=np.int64)
    stage = np.zeros(Nn, dtype=np.int64)
    real = np.full(Nn, -1, dtype=np.int64)
    nxt = 1
    for i in range(Nn):
        for j in range(nk[i]):
            k = nxt + j
            dad[k], stage[k] = i, stage[i] + 1
            real[k] = j if nk[i] > 1 else (real[i] if i > 0 else 0)
        nxt += nk[i]
    assert nxt == Nn
    return dad, stage, real


def spring_mass(md: int = 3, Nr: int = 2, Nh: int = 10, xmax1: float | None = None,
                x0=None) -> Problem:
    """Config 1/2: the reference's spring-mass example (examples/spring_mass_utils/data.c via
    tree_qp_in_fill_lti_data_diag_weights, tree_qp_common.c:1837-1955).  x0 is pinned through
    xmin[0]=xmax[0]=x0 as in examples/spring_mass_dual_newton_tree.c:105."""
    d = np.load(os.path.join(_GOLDEN, "spring_mass_data.npz"))
    NX, NU = int(d["NX"]), int(d["NU"])
    # first realisation in the file is the nominal one and is skipped (example :105)
    A = d["A"].reshape(-1, NX * NX)[1:]
    B = d["B"].reshape(-1, NX * NU)[1:]
    b = d["b"].reshape(-1, NX)[1:]
    assert md <= A.shape[0]
    nk = multistage_nk(md, Nr, Nh)
    Nn = len(nk)
    dad, stage, real = tree_tables(nk)
    nx = np.full(Nn, NX, dtype=np.int32)
    nu = np.where(nk > 0, NU, 0).astype(np.int32)
    x0 = d["x0"] if x0 is None else np.asarray(x0, dtype=float)
    xmax = d["xmax"].copy()
    if xmax1 is not None:
        xmax[1] = xmax1

    n_leaves = int((stage == stage[-1]).sum())
    As, Bs, bs, Qd, Rd, q, r, xlo, xhi, ulo, uhi = ([] for _ in range(11))
    for k in range(Nn):
        if k > 0:
            As.append(A[real[k]]); Bs.append(B[real[k]]); bs.append(b[real[k]])
        if nk[k] > 0:
            # integer division exactly as the reference (tree_qp_common.c:1909)
            scale = float(n_leaves // int((stage == stage[k]).sum()))
            Qd.append(scale * d["dQ"]); q.append(scale * d["q"])
            Rd.append(scale * d["dR"]); r.append(scale * d["r"])
            ulo.append(d["umin"]); uhi.append(d["umax"])
        else:
            Qd.append(d["dP"]); q.append(d["p"])
        if k == 0:
            xlo.append(x0); xhi.append(x0)
        else:
            xlo.append(d["xmin"]); xhi.append(xmax)
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    return Problem(nk=nk, nx=nx, nu=nu, A=cat(As), B=cat(Bs), b=cat(bs), Qd=cat(Qd), Rd=cat(Rd),
                   q=cat(q), r=cat(r), xmin=cat(xlo), xmax=cat(xhi), umin=cat(ulo), umax=cat(uhi),
                   lam0=np.zeros(int(nx.sum() - nx[0])),
                   meta=dict(name="spring_mass", md=md, Nr=Nr, Nh=Nh))


def random_tree(nk: np.ndarray, nx: int = 12, nu: int = 4, seed: int = 0, ubound: float = 0.5,
                xbound: float | None = None, a_scale: float | None = None,
                vary_dims: bool = False) -> Problem:
    """Config 3: random tree QP with diagonal weights (clipping path), restating the clipping
    branch of examples/random_qp_utils/generate_random_tree.m:52-79 with the three documented
    deviations of SURVEY.md §8d.3: A = a_scale*U(0,1) (default 0.96/nx), weights bounded away
    from zero, bounds on the inputs only (|u|<=ubound) unless xbound is given.

    vary_dims=True draws nx in [1,nx], nu in [0,nu] per node (the shape of the reference's unit
    test data, examples/random_qp_utils/data0x.c)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nk = np.asarray(nk, dtype=np.int32)
    Nn = len(nk)
    dad, stage, _ = tree_tables(nk)
    if vary_dims:
        nxs = rng.integers(1, nx + 1, size=Nn).astype(np.int32)
        nus = rng.integers(0, nu + 1, size=Nn).astype(np.int32)
    else:
        nxs = np.full(Nn, nx, dtype=np.int32)
        nus = np.full(Nn, nu, dtype=np.int32)
    nus[nk == 0] = 0
    if a_scale is None:
        a_scale = 0.96 / nx
    nA = int((nxs[1:].astype(np.int64) * nxs[dad[1:]]).sum())
    nB = int((nxs[1:].astype(np.int64) * nus[dad[1:]]).sum())
    X, U = int(nxs.sum()), int(nus.sum())
    A = a_scale * rng.random(nA)
    B = rng.random(nB)
    b = rng.random(X - int(nxs[0]))
    Qd = 1.0 + 9.0 * rng.random(X)
    Rd = 0.5 + rng.random(U)
    q = rng.random(X)
    r = rng.random(U)
    xb = TREEQP_INF if xbound is None else xbound
    return Problem(nk=nk, nx=nxs, nu=nus, A=A, B=B, b=b, Qd=Qd, Rd=Rd, q=q, r=r,
                   xmin=np.full(X, -xb), xmax=np.full(X, xb),
                   umin=np.full(U, -ubound), umax=np.full(U, ubound),
                   lam0=np.zeros(X - int(nxs[0])),
                   meta=dict(name="random_tree", seed=seed, nx=nx, nu=nu, ubound=ubound))


def densify(prob: Problem, seed: int = 1, coupling: float = 0.1, s_scale: float = 0.3) -> Problem:
    """Config 5: the same QP with genuinely dense stage Hessians, as the non-diagonal branch of
    examples/random_qp_utils/generate_random_tree.m:65-70 does: Q <- diag(Qd) + coupling*(G+G'),
    R likewise, S ~ s_scale*U(0,1) (nu x nx).  Needs the dense stage-QP path
    (qp_solver = TREEQP_QPOASES_SOLVER).  coupling = s_scale = 0 keeps the weights diagonal but
    still stores them as dense blocks."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Q, R, S = [], [], []
    ix = iu = 0
    for k in range(prob.Nn):
        nx, nu = int(prob.nx[k]), int(prob.nu[k])
        G = rng.random((nx, nx))
        Q.append((np.diag(prob.Qd[ix:ix + nx]) + coupling * (G + G.T)).flatten(order="F"))
        G = rng.random((nu, nu))
        R.append((np.diag(prob.Rd[iu:iu + nu]) + coupling * (G + G.T)).flatten(order="F"))
        S.append((s_scale * rng.random((nu, nx))).flatten(order="F"))
        ix += nx; iu += nu
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    return Problem(nk=prob.nk, nx=prob.nx, nu=prob.nu, A=prob.A, B=prob.B, b=prob.b, q=prob.q, r=prob.r,
                   xmin=prob.xmin, xmax=prob.xmax, umin=prob.umin, umax=prob.umax,
                   Q=cat(Q), R=cat(R), S=cat(S), lam0=prob.lam0, meta=dict(prob.meta, dense=True))


def binary_tree_nk(depth: int, md: int = 2) -> np.ndarray:
    """full md-ary tree with `depth` edges on every root-to-leaf path"""
    return multistage_nk(md, depth, depth)


def _expm(M: np.ndarray) -> np.ndarray:
    """scaling-and-squaring Taylor exponential (enough for Ts*A of a spring chain)"""
    nrm = np.linalg.norm(M, 1)
    s = max(0, int(np.ceil(np.log2(max(nrm, 1e-16)))) + 4)
    Ms = M / (2.0 ** s)
    E = np.eye(M.shape[0])
    term = np.eye(M.shape[0])
    for i in range(1, 20):
        term = term @ Ms / i
        E = E + term
    for _ in range(s):
        E = E @ E
    return E


def linear_chain(nm: int = 4, nu: int = 3, md: int = 3, Nr: int = 2, Nh: int = 20, seed: int = 0,
                 Ts: float = 0.05, eliminate_x0: bool = True, x0_kick: float = 1.5,
                 x0_noise: float = 0.3) -> Problem:
    """Config 4: chain of nm masses (nx = 2*nm), restating
    benchmark/linear_chain/initialize_linear_chain.m:31-81 + default_params_linear_chain.m:6-23:
    T = tridiag(1,-2,1), A_c=[0 I; k*T 0], B_c=[0; I_nu; 0], exact discretisation
    expm([A_c B_c; 0 0]*Ts) (benchmark/common/discretize_model.m:8-10), k in linspace(4,8,md);
    Q=10, R=1, P=Q, all bounds +-2, stage weights scaled md^(Nr-stage)
    (benchmark/utils/setup_tree.m:84-91); x0 = x0_kick*e_{nm+nu+1} + x0_noise*U(-1,1) (the benchmark
    starts at 2*e; 1.5 +- 0.3 keeps every instance of a batch strictly inside the +-2 box so
    that the QP is feasible for every seed), eliminated into the
    root's outgoing edges as tree_qp_in_eliminate_x0 does (tree_qp_common.c:404-525): nx[0]=0,
    b <- A0 x0 + b."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nxs_ = 2 * nm
    T = -2.0 * np.eye(nm) + np.eye(nm, k=1) + np.eye(nm, k=-1)
    Ad, Bd = [], []
    for kspr in np.linspace(4.0, 8.0, md):
        Ac = np.zeros((nxs_, nxs_))
        Ac[:nm, nm:] = np.eye(nm)
        Ac[nm:, :nm] = kspr * T
        Bc = np.zeros((nxs_, nu))
        Bc[nm:nm + nu, :] = np.eye(nu)
        M = np.zeros((nxs_ + nu, nxs_ + nu))
        M[:nxs_, :nxs_] = Ac
        M[:nxs_, nxs_:] = Bc
        E = _expm(M * Ts)
        Ad.append(E[:nxs_, :nxs_].copy())
        Bd.append(E[:nxs_, nxs_:].copy())
    nk = multistage_nk(md, Nr, Nh)
    Nn = len(nk)
    dad, stage, real = tree_tables(nk)
    x0 = np.zeros(nxs_)
    x0[min(nm + nu, nxs_ - 1)] = x0_kick
    x0 += x0_noise * rng.uniform(-1.0, 1.0, nxs_)
    nx = np.full(Nn, nxs_, dtype=np.int32)
    nus = np.where(nk > 0, nu, 0).astype(np.int32)
    if eliminate_x0:
        nx[0] = 0
    As, Bs, bs, Qd, Rd, q, r = ([] for _ in range(7))
    for k in range(Nn):
        w = float(md ** max(Nr - stage[k], 0))
        if k > 0:
            Ak, Bk = Ad[real[k]], Bd[real[k]]
            if dad[k] == 0 and eliminate_x0:
                As.append(np.zeros(0)); bs.append(Ak @ x0)
            else:
                As.append(Ak.flatten(order="F")); bs.append(np.zeros(nxs_))
            Bs.append(Bk.flatten(order="F"))
        if nx[k] > 0:
            Qd.append(np.full(nxs_, 10.0 * w)); q.append(np.zeros(nxs_))
        if nus[k] > 0:
            Rd.append(np.full(nu, 1.0 * w)); r.append(np.zeros(nu))
    X, U = int(nx.sum()), int(nus.sum())
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    p = Problem(nk=nk, nx=nx, nu=nus, A=cat(As), B=cat(Bs), b=cat(bs), Qd=cat(Qd), Rd=cat(Rd),
                q=cat(q), r=cat(r), xmin=np.full(X, -2.0), xmax=np.full(X, 2.0),
                umin=np.full(U, -2.0), umax=np.full(U, 2.0), lam0=np.zeros(X - int(nx[0])),
                meta=dict(name="linear_chain", nm=nm, nu=nu, md=md, Nr=Nr, Nh=Nh, seed=seed))
    if not eliminate_x0:
        p.xmin[:nxs_] = x0
        p.xmax[:nxs_] = x0
    return p


def linear_chain_batch(n: int, nm: int = 4, nu: int = 3, md: int = 3, Nr: int = 2, Nh: int = 20,
                       seed: int = 0, x0_kick: float = 1.0, x0_noise: float = 0.3, ubound: float = 0.5,
                       xbound: float = 1e12):
    """Config 4: n independent linear_chain MPC QPs of the same shape that differ only in the
    initial state (seed + i), i.e. only in b of the root's children (x0 is eliminated).
    Deviations from benchmark/linear_chain (all bounds +-2, x0 = 2*e), chosen on the reference
    itself: whenever a state reaches its bound the reference's TDUNES may not converge in 100
    iterations (x0 = 1.5 +- 0.3: 5 % of the seeds; x0 = 1.0 +- 0.3: 2 of 4096), so the batch bounds the
    inputs only, as config 3 does: |u| <= 0.5 makes ~250 of the 522 inputs active (3-5 Newton
    iterations per QP instead of a single unconstrained one) and the dual Hessian stays positive
    definite, i.e. every instance converges."""
    base = linear_chain(nm, nu, md, Nr, Nh, seed=seed, x0_kick=x0_kick, x0_noise=x0_noise)
    base.umin[:] = -ubound
    base.umax[:] = ubound
    base.xmin[:] = -xbound
    base.xmax[:] = xbound
    nxs_ = 2 * nm
    dad, stage, real = tree_tables(base.nk)
    kids0 = [k for k in range(1, base.Nn) if dad[k] == 0]
    Ts = 0.05
    T = -2.0 * np.eye(nm) + np.eye(nm, k=1) + np.eye(nm, k=-1)
    Ad = []
    for kspr in np.linspace(4.0, 8.0, md):
        Ac = np.zeros((nxs_, nxs_)); Ac[:nm, nm:] = np.eye(nm); Ac[nm:, :nm] = kspr * T
        Bc = np.zeros((nxs_, nu)); Bc[nm:nm + nu, :] = np.eye(nu)
        M = np.zeros((nxs_ + nu, nxs_ + nu)); M[:nxs_, :nxs_] = Ac; M[:nxs_, nxs_:] = Bc
        Ad.append(_expm(M * Ts)[:nxs_, :nxs_].copy())
    probs = []
    for i in range(n):
        rng = np.random.Generator(np.random.PCG64(seed + i))
        x0 = np.zeros(nxs_)
        x0[min(nm + nu, nxs_ - 1)] = x0_kick
        x0 += x0_noise * rng.uniform(-1.0, 1.0, nxs_)
        b = base.b.copy()
        for k in kids0:
            b[(k - 1) * nxs_:k * nxs_] = Ad[real[k]] @ x0
        probs.append(Problem(nk=base.nk, nx=base.nx, nu=base.nu, A=base.A, B=base.B, b=b, Qd=base.Qd, Rd=base.Rd,
                             q=base.q, r=base.r, xmin=base.xmin, xmax=base.xmax, umin=base.umin, umax=base.umax,
                             lam0=base.lam0, meta=dict(base.meta, seed=seed + i)))
    return probs


def eliminate_x0(prob: Problem) -> Problem:
    """Fold a pinned initial state (xmin[0] == xmax[0] == x0) into the dynamics of the root's
    children, as tree_qp_in_eliminate_x0 does (tree_qp_common.c:404-525): nx[0] <- 0,
    b_k <- A_k x0 + b_k for the children k of the root, A_k loses its columns.  This is the form
    SDUNES requires (dual_Newton_scenarios.c:1498) and the JSON driver always uses
    (solve_qp_json.cpp:353)."""
    nx0 = int(prob.nx[0])
    if nx0 == 0:
        return prob
    x0 = np.asarray(prob.xmin[:nx0], dtype=float)
    assert np.array_equal(x0, prob.xmax[:nx0]), "x0 is not pinned"
    dad = prob.dad
    As, bs = [], []
    ia = ix = 0
    for k in range(1, prob.Nn):
        nxk, nxp = int(prob.nx[k]), int(prob.nx[dad[k]])
        Ak = prob.A[ia:ia + nxk * nxp]
        bk = prob.b[ix:ix + nxk]
        if dad[k] == 0:
            bs.append(Ak.reshape(nxp, nxk).T @ x0 + bk)
        else:
            As.append(Ak); bs.append(bk)
        ia += nxk * nxp; ix += nxk
    nx = prob.nx.copy(); nx[0] = 0
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    return Problem(nk=prob.nk, nx=nx, nu=prob.nu, A=cat(As), B=prob.B, b=cat(bs), Qd=prob.Qd[nx0:],
                   Rd=prob.Rd, q=prob.q[nx0:], r=prob.r, xmin=prob.xmin[nx0:], xmax=prob.xmax[nx0:],
                   umin=prob.umin, umax=prob.umax, lam0=prob.lam0,
                   meta=dict(prob.meta, x0_eliminated=True))


def scenario_dims(prob: Problem):
    """(Ns, Nh, Nr, md, nl): scenarios, horizon, robust horizon, branching and the number of
    non-anticipativity multipliers (treeqp_sdunes_calculate_dual_dimension,
    dual_Newton_scenarios.c:99-107)."""
    nk = np.asarray(prob.nk)
    _, stage, _ = tree_tables(nk)
    Ns, Nh, md = int((nk == 0).sum()), int(stage.max()), int(nk[0])
    Nr = 0
    for k in range(len(nk)):
        if nk[k] > 1:
            Nr = int(stage[k]) + 1
        else:
            break
    nu = int(prob.nu[0])
    nl = -1 if Ns == 1 else (Nr * Ns - (Ns - 1) // (md - 1)) * nu
    return Ns, Nh, Nr, md, nl


# ------------------------------------------------------------------ benchmark models of config 5
def _benchmark_tree(Ad, Bd, Qw, Rw, Pw, xref, uref, xmin, xmax, umin, umax, xNmin, xNmax, x0,
                    md, Nr, Nh, name, eliminate: bool = True) -> Problem:
    """Tree QP of the benchmark harness (benchmark/utils/setup_tree.m): realisation j of the dynamics on
    the j-th child of every branching node, stage weights scaled by md^(Nr - stage) up to the robust
    horizon (:84-91; the root gets Nscenarios = md^Nr, :48-51), terminal weight P unscaled on the leaves
    (:97-100), linear terms q = -Q.*xref, r = -R.*uref, p = q (benchmark/treeqp_main.m:147-160), leaf
    bounds xNmin/xNmax, x0 pinned on the root and eliminated like the JSON driver does
    (examples/solve_qp_json.cpp:353)."""
    nxs_, nu = Bd[0].shape
    nk = multistage_nk(md, Nr, Nh)
    Nn = len(nk)
    dad, stage, real = tree_tables(nk)
    nx = np.full(Nn, nxs_, dtype=np.int32)
    nus = np.where(nk > 0, nu, 0).astype(np.int32)
    As, Bs, bs, Qd, Rd, q, r, lo, hi, ulo, uhi = ([] for _ in range(11))
    for k in range(Nn):
        leaf = nk[k] == 0
        w = 1.0 if leaf else float(md ** max(Nr - stage[k], 0))
        if k > 0:
            As.append(Ad[real[k]].flatten(order="F")); Bs.append(Bd[real[k]].flatten(order="F")); bs.append(np.zeros(nxs_))
        Qk = Pw if leaf else Qw
        Qd.append(w * Qk); q.append(-(w * Qk) * xref)
        lo.append(xNmin if leaf else xmin); hi.append(xNmax if leaf else xmax)
        if not leaf:
            Rd.append(w * Rw); r.append(-(w * Rw) * uref); ulo.append(umin); uhi.append(umax)
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    p = Problem(nk=nk, nx=nx, nu=nus, A=cat(As), B=cat(Bs), b=cat(bs), Qd=cat(Qd), Rd=cat(Rd), q=cat(q), r=cat(r),
                xmin=cat(lo), xmax=cat(hi), umin=cat(ulo), umax=cat(uhi), lam0=np.zeros(int(nx.sum()) - nxs_),
                meta=dict(name=name, md=md, Nr=Nr, Nh=Nh))
    p.xmin[:nxs_] = x0
    p.xmax[:nxs_] = x0
    return eliminate_x0(p) if eliminate else p


def _discretize(Ac, Bc, Ts):
    """benchmark/common/discretize_model.m:8-10: expm([Ts*A Ts*B; 0 0])"""
    n, m = Bc.shape
    M = np.zeros((n + m, n + m))
    M[:n, :n] = Ts * Ac
    M[:n, n:] = Ts * Bc
    E = _expm(M)
    return E[:n, :n].copy(), E[:n, n:].copy()


def quadcopter(md: int = 3, Nr: int = 2, Nh: int = 20, i_mpc: int = 12, nsim: int = 50, x0=None) -> Problem:
    """Config 5, quadcopter attitude model (nx = 6: quaternion vector part + body rates, nu = 4 rotor speed
    deviations), restating benchmark/quadcopter/initialize_quadcopter.m:33-114:
    dynamics_quadcopter_mpc.m linearised at hover (x = 0, W_i = omega_hover(m)) -- the Jacobians the
    harness gets from CasADi, written out by hand: d(quat)/d(Omega) = I/2, dOmega1/dW = a L [0 1 0 -1]/J1,
    dOmega2/dW = a L [-1 0 1 0]/J2, dOmega3/dW = a_d L2 [1 -1 1 -1]/J3 with a = A Cl rho omega_hover,
    a_d = A Cd rho omega_hover -- discretised exactly (Ts = 0.05), one realisation per mass in
    linspace(8, 12, md) (import_benchmark.m:52-59); Q = [500 500 500 1e-3 1e-3 1e-3], R = 1e-3, P = Q,
    |Omega| <= 1, |dW| <= 4, terminal state free, x0 = 0; reference = sample i_mpc of the attitude
    reference (:57-83, myeul2quat.m)."""
    rho, Ar, Cl, g, L, L2, J1, J2, J3, inf, Ts = 1.23, 0.1, 0.25, 9.81, 0.5, 1.0, 0.25, 0.25, 1.0, 1e8, 0.05
    Cd = 0.3 * Cl
    Ad, Bd = [], []
    for m in np.linspace(8.0, 12.0, md):
        wh = np.sqrt(2.0 * m * g / (Ar * Cl * rho) / 4.0)
        Ac = np.zeros((6, 6)); Ac[:3, 3:] = 0.5 * np.eye(3)
        Bc = np.zeros((6, 4))
        a, ad = Ar * Cl * rho * wh, Ar * Cd * rho * wh
        Bc[3] = a * L * np.array([0.0, 1.0, 0.0, -1.0]) / J1
        Bc[4] = a * L * np.array([-1.0, 0.0, 1.0, 0.0]) / J2
        Bc[5] = ad * L2 * np.array([1.0, -1.0, 1.0, -1.0]) / J3
        A_, B_ = _discretize(Ac, Bc, Ts)
        Ad.append(A_); Bd.append(B_)
    # attitude reference (initialize_quadcopter.m:57-83)
    period, ang, lam = nsim // 4, -1.0 * np.pi / 2 / 9 * 5 * 0.1, 0.3
    ref = np.zeros(3)
    for i in range(i_mpc):
        ph = (i // period) % 3
        s1, s2 = (+1, -1) if ph == 0 else ((-1, -1) if ph == 1 else (-1, +1))
        ref = np.array([ref[0] - lam * (ref[0] + s1 * ang), ref[1] - lam * (ref[1] + s2 * ang), 0.0])
    rr, pp, yy = ref
    cy, sy, cr, sr, cp, sp = np.cos(yy / 2), np.sin(yy / 2), np.cos(rr / 2), np.sin(rr / 2), np.cos(pp / 2), np.sin(pp / 2)
    xref = np.array([cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp, 0.0, 0.0, 0.0])
    Qw = np.array([500.0, 500.0, 500.0, 1e-3, 1e-3, 1e-3])
    xmax = np.array([inf, inf, inf, 1.0, 1.0, 1.0])
    return _benchmark_tree(Ad, Bd, Qw, np.full(4, 1e-3), Qw, xref, np.zeros(4), -xmax, xmax, np.full(4, -4.0), np.full(4, 4.0),
                           np.full(6, -inf), np.full(6, inf), np.zeros(6) if x0 is None else np.asarray(x0, float),
                           md, Nr, Nh, "quadcopter")


def crane(md: int = 3, Nr: int = 2, Nh: int = 20, xref0: float = 0.2, x0=None) -> Problem:
    """Config 5, overhead crane (nx = 4: cart position/velocity, cable angle/rate; nu = 1 cart
    acceleration), restating benchmark/crane/initialize_crane.m:33-72 with dynamics_crane.m linearised at
    the origin: A_c = [0 1 0 0; 0 0 0 0; 0 0 0 1; 0 0 -g -b], B_c = [0; 1; 0; -1], Ts = 0.2, friction
    b in linspace(0.1, 0.3, md) (import_benchmark.m:72-79); Q = [10 1 1 1], R = 0.1, P = Q, |v| <= 0.2,
    |omega| <= 0.4 (also on the leaves), |a| <= 0.5, x0 = 0, position reference 0.2."""
    g, inf, Ts = 9.81, 1e8, 0.2
    Ad, Bd = [], []
    for bf in np.linspace(0.1, 0.3, md):
        Ac = np.zeros((4, 4)); Ac[0, 1] = 1.0; Ac[2, 3] = 1.0; Ac[3, 2] = -g; Ac[3, 3] = -bf
        Bc = np.array([[0.0], [1.0], [0.0], [-1.0]])
        A_, B_ = _discretize(Ac, Bc, Ts)
        Ad.append(A_); Bd.append(B_)
    Qw = np.array([10.0, 1.0, 1.0, 1.0])
    xmax = np.array([inf, 0.2, inf, 0.4])
    return _benchmark_tree(Ad, Bd, Qw, np.array([0.1]), Qw, np.array([xref0, 0.0, 0.0, 0.0]), np.zeros(1), -xmax, xmax,
                           np.array([-0.5]), np.array([0.5]), -xmax, xmax, np.zeros(4) if x0 is None else np.asarray(x0, float),
                           md, Nr, Nh, "crane")


def densify_scaled(prob: Problem, seed: int = 1, coupling: float = 0.1, s_scale: float = 0.05) -> Problem:
    """Dense stage Hessians for weights that span decades (quadcopter: 500 ... 1e-3): the coupling of
    examples/random_qp_utils/generate_random_tree.m:65-70 applied in the scaled metric,
    Q <- D^(1/2) (I + coupling (G + G')/n) D^(1/2) with D = diag(Qd) (positive definite for coupling < 1),
    R likewise, S <- s_scale * sqrt(R_i Q_j) * U(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Q, R, S = [], [], []
    ix = iu = 0
    for k in range(prob.Nn):
        nx, nu = int(prob.nx[k]), int(prob.nu[k])
        dq, dr = np.sqrt(prob.Qd[ix:ix + nx]), np.sqrt(prob.Rd[iu:iu + nu])
        G = rng.random((nx, nx))
        Q.append((np.outer(dq, dq) * (np.eye(nx) + coupling * (G + G.T) / max(nx, 1))).flatten(order="F"))
        G = rng.random((nu, nu))
        R.append((np.outer(dr, dr) * (np.eye(nu) + coupling * (G + G.T) / max(nu, 1))).flatten(order="F"))
        S.append((s_scale * np.outer(dr, dq) * rng.random((nu, nx))).flatten(order="F"))
        ix += nx; iu += nu
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    return Problem(nk=prob.nk, nx=prob.nx, nu=prob.nu, A=prob.A, B=prob.B, b=prob.b, q=prob.q, r=prob.r,
                   xmin=prob.xmin, xmax=prob.xmax, umin=prob.umin, umax=prob.umax,
                   Q=cat(Q), R=cat(R), S=cat(S), lam0=prob.lam0, meta=dict(prob.meta, dense=True))


# ------------------------------------------------------------------ general stage constraints
def bounds_to_general(prob: Problem, xidx=(1,), uidx=(0,)) -> Problem:
    """The reference's own general-constraint test data (examples/spring_mass.c:100-223, TEST_GENERAL_CONSTRAINTS):
    the bounds on x[xidx] and u[uidx] of every node become rows of dmin <= C x + D u <= dmax (unit rows) and
    the bounds themselves are removed; a pinned root state keeps its bounds.  Same QP, same solution."""
    import copy
    p = copy.deepcopy(prob)
    nc, Cl, Dl, lo_l, hi_l = [], [], [], [], []
    ix = iu = 0
    for k in range(p.Nn):
        nx, nu = int(p.nx[k]), int(p.nu[k])
        rows = [("x", j) for j in xidx if j < nx and not (k == 0 and p.xmin[ix + j] == p.xmax[ix + j])]
        rows += [("u", j) for j in uidx if j < nu]
        n = len(rows)
        Ck, Dk, lo, hi = np.zeros((n, nx)), np.zeros((n, nu)), np.zeros(n), np.zeros(n)
        for i, (kind, j) in enumerate(rows):
            if kind == "x":
                Ck[i, j] = 1.0; lo[i], hi[i] = p.xmin[ix + j], p.xmax[ix + j]
                p.xmin[ix + j], p.xmax[ix + j] = -TREEQP_INF, TREEQP_INF
            else:
                Dk[i, j] = 1.0; lo[i], hi[i] = p.umin[iu + j], p.umax[iu + j]
                p.umin[iu + j], p.umax[iu + j] = -TREEQP_INF, TREEQP_INF
        nc.append(n); Cl.append(Ck.flatten(order="F")); Dl.append(Dk.flatten(order="F")); lo_l.append(lo); hi_l.append(hi)
        ix += nx; iu += nu
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    p.nc, p.C, p.D, p.dmin, p.dmax = np.asarray(nc, dtype=np.int32), cat(Cl), cat(Dl), cat(lo_l), cat(hi_l)
    p.meta = dict(p.meta, general="bounds_to_general")
    return p


def add_general_rows(prob: Problem, seed: int = 1, nrows: int = 2, width: float = 0.3, leaf_rows: int = 1) -> Problem:
    """Random coupled rows -width <= c'x + d'u <= width (c, d ~ U(-1, 1), so the origin is strictly feasible)
    on every node: `nrows` on parents, `leaf_rows` on leaves."""
    import copy
    p = copy.deepcopy(prob)
    rng = np.random.Generator(np.random.PCG64(seed))
    nc, Cl, Dl, lo_l, hi_l = [], [], [], [], []
    for k in range(p.Nn):
        nx, nu = int(p.nx[k]), int(p.nu[k])
        n = nrows if nu > 0 else leaf_rows
        if nx == 0:
            n = min(n, 1) if nu > 0 else 0
        Ck, Dk = rng.uniform(-1.0, 1.0, (n, nx)), rng.uniform(-1.0, 1.0, (n, nu))
        nc.append(n); Cl.append(Ck.flatten(order="F")); Dl.append(Dk.flatten(order="F"))
        lo_l.append(np.full(n, -width)); hi_l.append(np.full(n, width))
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    p.nc, p.C, p.D, p.dmin, p.dmax = np.asarray(nc, dtype=np.int32), cat(Cl), cat(Dl), cat(lo_l), cat(hi_l)
    p.meta = dict(p.meta, general="add_general_rows", general_seed=seed)
    return p

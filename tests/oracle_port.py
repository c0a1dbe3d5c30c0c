"""ctypes wrapper of oracle/_ref/liboracle_port.so (the plain-C restatement, oracle/tdunes_port.c).
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PORT_LIB = os.path.join(ROOT, "oracle", "_ref", "liboracle_port.so")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libtreeqp_ref.so")
REF_LIB_OMP = os.path.join(ROOT, "oracle", "_ref", "libtreeqp_ref_omp.so")

_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)


class PortQp(C.Structure):
    _fields_ = [("Nn", C.c_int), ("nk", _ip), ("nx", _ip), ("nu", _ip), ("A", _dp), ("B", _dp), ("b", _dp),
                ("Qd", _dp), ("Rd", _dp), ("q", _dp), ("r", _dp), ("xmin", _dp), ("xmax", _dp),
                ("umin", _dp), ("umax", _dp), ("solver", _ip), ("H", _dp)]


class PortOpts(C.Structure):
    _fields_ = [("maxIter", C.c_int), ("termCondition", C.c_int), ("tol", C.c_double), ("regType", C.c_int),
                ("regTol", C.c_double), ("regValue", C.c_double), ("lsMaxIter", C.c_int),
                ("gamma", C.c_double), ("beta", C.c_double)]


def build_port():
    """compile the port if missing (gcc only; needs nothing from /root/reference)"""
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("tdunes_port.c", "sdunes_port.c", "port_common.h")]
    if not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return PORT_LIB


def dense_blocks(prob):
    """per node H = [Q S'; S R] ((nx+nu)^2 column-major, concatenated) and the diagonals of Q, R"""
    H, Qd, Rd = [], [], []
    iq = ir = is_ = 0
    for k in range(prob.Nn):
        nx, nu = int(prob.nx[k]), int(prob.nu[k])
        Q = np.asarray(prob.Q[iq:iq + nx * nx]).reshape(nx, nx, order="F")
        R = np.asarray(prob.R[ir:ir + nu * nu]).reshape(nu, nu, order="F")
        S = np.asarray(prob.S[is_:is_ + nu * nx]).reshape(nu, nx, order="F")
        Hk = np.block([[Q, S.T], [S, R]]) if nu else Q
        H.append(Hk.flatten(order="F")); Qd.append(np.diag(Q).copy()); Rd.append(np.diag(R).copy())
        iq += nx * nx; ir += nu * nu; is_ += nu * nx
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    return cat(H), cat(Qd), cat(Rd)


def solve_port(prob, lam0=None, qp_solver=None, maxIter=100, termCondition=2, tol=1e-8, regType=2, regTol=1e-6,
               regValue=1e-6, lineSearchMaxIter=50, lineSearchGamma=0.1, lineSearchBeta=0.6):
    lib = C.CDLL(build_port())
    f = lib.oracle_tdunes_solve
    f.restype = C.c_int
    keep = {}

    def d(a):
        a = np.ascontiguousarray(a, dtype=np.float64); keep[id(a)] = a
        return a.ctypes.data_as(_dp)

    def i(a):
        a = np.ascontiguousarray(a, dtype=np.int32); keep[id(a)] = a
        return a.ctypes.data_as(_ip)

    Qd, Rd, solver, H = prob.Qd, prob.Rd, None, None
    if prob.Qd is None:
        H, Qd, Rd = dense_blocks(prob)
        assert qp_solver is not None, "dense weights need the dense stage solver"
    if qp_solver is not None:
        solver = i(np.full(prob.Nn, qp_solver) if np.isscalar(qp_solver) else qp_solver)
        if H is None:   # diagonal weights through the dense path
            H = np.concatenate([np.diag(np.concatenate([prob.Qd[xo:xo + nx], prob.Rd[uo:uo + nu]])).flatten()
                                for xo, nx, uo, nu in zip(np.cumsum(prob.nx) - prob.nx, prob.nx,
                                                          np.cumsum(prob.nu) - prob.nu, prob.nu)])
        H = d(H)
    qp = PortQp(prob.Nn, i(prob.nk), i(prob.nx), i(prob.nu), d(prob.A), d(prob.B), d(prob.b), d(Qd),
                d(Rd), d(prob.q), d(prob.r), d(prob.xmin), d(prob.xmax), d(prob.umin), d(prob.umax), solver, H)
    o = PortOpts(maxIter, termCondition, tol, regType, regTol, regValue, lineSearchMaxIter,
                 lineSearchGamma, lineSearchBeta)
    X, U = int(np.sum(prob.nx)), int(np.sum(prob.nu))
    nl = X - int(prob.nx[0])
    lam0 = np.zeros(nl) if lam0 is None else np.ascontiguousarray(lam0, dtype=np.float64)
    x, u, lam = np.zeros(X + 1), np.zeros(U + 1), np.zeros(nl + 1)
    mux, muu = np.zeros(X + 1), np.zeros(U + 1)
    ls = np.zeros(maxIter + 1, dtype=np.int32)
    xas, uas = np.zeros(X + 1, dtype=np.int32), np.zeros(U + 1, dtype=np.int32)
    it, nreg = C.c_int(0), C.c_int(0)
    status = f(C.byref(qp), C.byref(o), lam0.ctypes.data_as(_dp), x.ctypes.data_as(_dp), u.ctypes.data_as(_dp),
               lam.ctypes.data_as(_dp), mux.ctypes.data_as(_dp), muu.ctypes.data_as(_dp), C.byref(it),
               ls.ctypes.data_as(_ip), xas.ctypes.data_as(_ip), uas.ctypes.data_as(_ip), C.byref(nreg))
    return dict(status=status, iter=it.value, x=x[:X], u=u[:U], lam=lam[:nl], mu_x=mux[:X], mu_u=muu[:U],
                ls_iters=ls[:it.value].tolist(), xas=xas[:X], uas=uas[:U], n_reg=nreg.value)


def solve_sdunes_port(prob, lam0=None, mu0=None, checkLastActiveSet=1, maxIter=100, termCondition=2, tol=1e-8,
                      regType=2, regTol=1e-6, regValue=1e-6, lineSearchMaxIter=50, lineSearchGamma=0.1,
                      lineSearchBeta=0.6):
    """oracle/sdunes_port.c on a Problem with x0 eliminated (nx[0] == 0)"""
    lib = C.CDLL(build_port())
    f = lib.oracle_sdunes_solve
    f.restype = C.c_int
    keep = {}

    def d(a):
        a = np.ascontiguousarray(a, dtype=np.float64); keep[id(a)] = a
        return a.ctypes.data_as(_dp)

    def i(a):
        a = np.ascontiguousarray(a, dtype=np.int32); keep[id(a)] = a
        return a.ctypes.data_as(_ip)

    qp = PortQp(prob.Nn, i(prob.nk), i(prob.nx), i(prob.nu), d(prob.A), d(prob.B), d(prob.b), d(prob.Qd),
                d(prob.Rd), d(prob.q), d(prob.r), d(prob.xmin), d(prob.xmax), d(prob.umin), d(prob.umax), None, None)
    o = PortOpts(maxIter, termCondition, tol, regType, regTol, regValue, lineSearchMaxIter,
                 lineSearchGamma, lineSearchBeta)
    Ns, Nh, Nr, nl = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib.oracle_sdunes_dims(C.byref(qp), C.byref(Ns), C.byref(Nh), C.byref(Nr), C.byref(nl))
    Ns, Nh, Nr, nl = Ns.value, Nh.value, Nr.value, nl.value
    nmu = Ns * Nh * int(prob.nx[1])
    X, U = int(np.sum(prob.nx)), int(np.sum(prob.nu))
    lam0 = np.zeros(nl + 1) if lam0 is None else np.ascontiguousarray(lam0, dtype=np.float64)
    mu0 = np.zeros(nmu + 1) if mu0 is None else np.ascontiguousarray(mu0, dtype=np.float64)
    x, u, lam = np.zeros(X + 1), np.zeros(U + 1), np.zeros(X + 1)
    mux, muu = np.zeros(X + 1), np.zeros(U + 1)
    lam_s, mu_s = np.zeros(nl + 1), np.zeros(nmu + 1)
    ls = np.zeros(maxIter + 1, dtype=np.int32)
    it, nreg = C.c_int(0), C.c_int(0)
    p = lambda a: a.ctypes.data_as(_dp)
    status = f(C.byref(qp), C.byref(o), C.c_int(checkLastActiveSet), p(lam0), p(mu0), p(x), p(u), p(lam), p(mux),
               p(muu), p(lam_s), p(mu_s), C.byref(it), ls.ctypes.data_as(_ip), C.byref(nreg))
    return dict(status=status, iter=it.value, x=x[:X], u=u[:U], lam=lam[:X - int(prob.nx[0])], mu_x=mux[:X],
                mu_u=muu[:U], ls_iters=ls[:it.value].tolist(), n_reg=nreg.value, lam_scen=lam_s[:nl],
                mu_scen=mu_s[:nmu], Ns=Ns, Nh=Nh, Nr=Nr)

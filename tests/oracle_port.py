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
                ("umin", _dp), ("umax", _dp)]


class PortOpts(C.Structure):
    _fields_ = [("maxIter", C.c_int), ("termCondition", C.c_int), ("tol", C.c_double), ("regType", C.c_int),
                ("regTol", C.c_double), ("regValue", C.c_double), ("lsMaxIter", C.c_int),
                ("gamma", C.c_double), ("beta", C.c_double)]


def build_port():
    """compile the port if missing (gcc only; needs nothing from /root/reference)"""
    if not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(
            os.path.join(ROOT, "oracle", "tdunes_port.c")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"])
    return PORT_LIB


def solve_port(prob, lam0=None, maxIter=100, termCondition=2, tol=1e-8, regType=2, regTol=1e-6,
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

    qp = PortQp(prob.Nn, i(prob.nk), i(prob.nx), i(prob.nu), d(prob.A), d(prob.B), d(prob.b), d(prob.Qd),
                d(prob.Rd), d(prob.q), d(prob.r), d(prob.xmin), d(prob.xmax), d(prob.umin), d(prob.umax))
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

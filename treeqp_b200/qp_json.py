"""The JSON wire format of the reference's benchmark harness (examples/solve_qp_json.cpp:206-615,
written on the MATLAB side by examples/random_qp_utils/code_generate_json.m and
benchmark/utils/treeqp_solve.m): Problem <-> JSON, and a runner for tools/solve_qp_json.cpp."""
import json
import os
import subprocess

import numpy as np

from .capi import Problem

EXE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "solve_qp_json")


def _mat(a, m, n):
    """what MATLAB's jsonencode emits: [] / number / flat array / array of rows"""
    a = np.asarray(a, dtype=np.float64).reshape(m, n, order="F")
    if m == 0 or n == 0:
        return []
    if m == 1 and n == 1:
        return float(a[0, 0])
    if m == 1 or n == 1:
        return [float(v) for v in a.flatten(order="F")]
    return [[float(v) for v in row] for row in a]


def _vec(a):
    a = np.asarray(a, dtype=np.float64)
    return [] if a.size == 0 else (float(a[0]) if a.size == 1 else [float(v) for v in a])


def problem_to_json(prob: Problem, options=None, init=None) -> dict:
    """nodes{Q,R,S,q,r,lx,ux,lu,uu}, edges{A,B,b,from,to}, options{...}, init{...}"""
    dad = prob.dad
    nodes, edges = [], []
    ix = iu = iq = ir = is_ = 0
    for k in range(prob.Nn):
        nx, nu = int(prob.nx[k]), int(prob.nu[k])
        if prob.Qd is not None:
            Q, R, S = np.diag(prob.Qd[ix:ix + nx]), np.diag(prob.Rd[iu:iu + nu]), np.zeros((nu, nx))
        else:
            Q = prob.Q[iq:iq + nx * nx].reshape(nx, nx, order="F")
            R = prob.R[ir:ir + nu * nu].reshape(nu, nu, order="F")
            S = prob.S[is_:is_ + nu * nx].reshape(nu, nx, order="F")
        nodes.append({"Q": _mat(Q.flatten(order="F"), nx, nx), "R": _mat(R.flatten(order="F"), nu, nu),
                      "S": _mat(S.flatten(order="F"), nu, nx), "q": _vec(prob.q[ix:ix + nx]), "r": _vec(prob.r[iu:iu + nu]),
                      "lx": _vec(prob.xmin[ix:ix + nx]), "ux": _vec(prob.xmax[ix:ix + nx]),
                      "lu": _vec(prob.umin[iu:iu + nu]), "uu": _vec(prob.umax[iu:iu + nu])})
        ix += nx; iu += nu; iq += nx * nx; ir += nu * nu; is_ += nu * nx
    ia = ib = iv = 0
    for k in range(1, prob.Nn):
        p, nxk = int(dad[k]), int(prob.nx[k])
        nxp, nup = int(prob.nx[p]), int(prob.nu[p])
        edges.append({"A": _mat(prob.A[ia:ia + nxk * nxp], nxk, nxp), "B": _mat(prob.B[ib:ib + nxk * nup], nxk, nup),
                      "b": _vec(prob.b[iv:iv + nxk]), "from": p, "to": k})
        ia += nxk * nxp; ib += nxk * nup; iv += nxk
    out = {"nodes": nodes, "edges": edges}
    if options:
        out["options"] = options
    if init:
        out["init"] = {k: (_vec(v) if not np.isscalar(v) else v) for k, v in init.items()}
    return out


def run_exe(problem_json: dict, workdir: str, extra_args=()) -> dict:
    """write the JSON, run tools/solve_qp_json (built by __graft_entry__.build), parse its output"""
    if not os.path.exists(EXE):
        raise FileNotFoundError(EXE + " (needs a nlohmann/json.hpp at build time)")
    path = os.path.join(workdir, "problem.json")
    with open(path, "w") as f:
        json.dump(problem_json, f)
    r = subprocess.run([EXE, path, *extra_args], capture_output=True, text=True)
    if r.returncode not in (0, 1):
        raise RuntimeError("solve_qp_json failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
    return json.loads(r.stdout)


def solution_arrays(res: dict):
    """concatenated x, u, lam of a result (node 0's x = the initial state when it was eliminated)"""
    cat = lambda key, items: np.concatenate([np.atleast_1d(np.asarray(i[key], dtype=np.float64)).ravel() for i in items]) if items else np.zeros(0)
    return cat("x", res["solution"]["nodes"]), cat("u", res["solution"]["nodes"]), cat("lam", res["solution"]["edges"])

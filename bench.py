#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native TDUNES hot path.

Metric (BASELINE.json): tree-QP solves/sec (and us per dual Newton iteration) on the 10^5-node
random tree (config 3: full binary tree of depth 16 = 131071 nodes, nx=12, nu=4, clipping stage
QPs, FP64).  One "step" = one complete dual Newton solve of that QP from the dual start lambda=0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--depth D]

  value  whole-job throughput with the QP data already resident in HBM (treeqp_tdunes_solve_resident)
  e2e    same metric through the drop-in C-ABI call treeqp_tdunes_solve with HOST buffers
         (BLASFEO views -> pinned SoA -> H2D -> Newton loop -> D2H -> BLASFEO views), every step
  roofline      dominant kernel class: algorithmic bytes (SURVEY.md 8d formulas) / CUDA-event time
  cpu_baseline  the unmodified reference (oracle/_ref, OpenMP build, all host cores) on the same QP

N>1: one process per GPU (torchrun), each rank solves its own copy of the same tree QP: the path
partitions by independent QP, no data-path collective -> weak scaling.  Next to it the record carries
ONE tree sharded by subtree over the N GPUs (strong scaling, NCCL exchange of the cut-node Schur
complements) and the 4096-QP batch split over the ranks.
--impl reference: times the reference's own CPU implementation (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from treeqp_b200 import capi, problems as P  # noqa: E402

METRIC = "tree_qp_solves_per_sec"
UNIT = "solves/s"


def workload(depth, seed):
    return P.random_tree(P.binary_tree_nk(depth), 12, 4, seed=seed, ubound=0.2)


OPTS = dict(maxIter=200)


# ------------------------------------------------------------------ algorithmic bytes
def algorithmic_bytes(prob):
    """bytes each kernel class must move per call, from the problem dimensions only"""
    nx, nu, nk = prob.nx.astype(np.int64), prob.nu.astype(np.int64), prob.nk.astype(np.int64)
    dad = prob.dad
    Nn = prob.Nn
    ab = np.zeros(Nn, dtype=np.int64)
    ab[1:] = nx[1:] * (nx[dad[1:]] + nu[dad[1:]])
    kid_ab = np.bincount(dad[1:], weights=ab[1:], minlength=Nn).astype(np.int64)
    d = np.bincount(dad[1:], weights=nx[1:], minlength=Nn).astype(np.int64)
    par = nk > 0
    nonroot_par = par.copy(); nonroot_par[0] = False
    k1 = 8 * int(((5 * nx + 4 * nu) + kid_ab + 2 * d + 5 * nx + 5 * nu + 1).sum())   # k_stage<0>
    k5 = 8 * int((2 * nx + kid_ab + 2 * d + nx + nu + 2).sum())                        # k_stage<2>
    trial = 8 * int(6 * nx.sum() + 6 * nu.sum() + 2 * Nn)                               # k_trial
    grad = 8 * int((ab[1:] + 4 * nx[1:] + nx[dad[1:]] + nu[dad[1:]]).sum())
    hess = 8 * int((kid_ab + nx + nu + d + d * (d + 1) // 2)[par].sum() + (nx * d)[nonroot_par].sum())
    k3 = 8 * int((d * (d + 1) + 2 * d)[par].sum() + (2 * nx * d + nx * (nx + 1) + 2 * nx)[nonroot_par].sum())
    k4 = 8 * int((d * (d + 1) // 2 + 2 * d)[par].sum() + (nx * d + nx)[nonroot_par].sum())
    return {"k_stage<0>": k1, "k_stage<2>": k5, "k_trial": trial, "k_gradient": grad, "k_hessian": hess,
            "k_factor_level": k3, "k_backsolve_level": k4}


def launches_per_call(prob):
    """how many launches make up ONE call of each phase (level kernels: one per depth)"""
    stage = P.tree_tables(prob.nk)[1]
    Nh = int(stage.max())
    return {"k_stage<0>": 1, "k_stage<2>": 1, "k_trial": 1, "k_gradient": 1, "k_hessian": 1,
            "k_factor_level": Nh, "k_backsolve_level": max(Nh - 1, 1)}


# ------------------------------------------------------------------------------ clock sampling
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a
    thread (nvidia-smi -lms as the fallback; it needs ~0.3 s to start, longer than a short run)."""
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
             "clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.on = index, [], None, None, False
        self.source = None

    def _nvml_loop(self):
        n, h = self.nvml, self.handle
        bits = [(n.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (n.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (n.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (n.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        while self.on:
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), self.mx, [nm for b, nm in bits if r & b]))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as n
            n.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    pass
            self.handle = n.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
            self.nvml, self.on, self.source = n, True, "nvml 5 ms"
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 20"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.rows.append((float(r[0]), float(r[1]), [n for n, v in zip(self.NAMES, r[3:7]) if v.lower().startswith("active")]))
            except Exception:
                pass

    def mark(self):
        """index of the next sample: brackets the timed region"""
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if self.nvml is not None:
            self.on = False
            self.thread.join(timeout=1)
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        rows = self.rows[lo:hi]
        sm = [r[0] for r in rows]
        mx = [r[1] for r in rows]
        reasons = sorted({n for r in rows for n in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


# ------------------------------------------------------------------------------ reference arm
def ref_lib_path(omp=True):
    p = os.path.join(ROOT, "oracle", "_ref", "libtreeqp_ref_omp.so" if omp else "libtreeqp_ref.so")
    return p if os.path.exists(p) else None


def run_reference_solve(prob, max_iter=None):
    """one treeqp_tdunes_solve of the UNMODIFIED reference (OpenMP build, all host cores)"""
    path = ref_lib_path(True)
    lib = capi.TreeQpLib(path)
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    opts = dict(OPTS)
    if max_iter is not None:
        opts["maxIter"] = max_iter
    solver = capi.Tdunes(lib, qp_in, **opts)
    lam0 = np.zeros(prob.n_lambda)

    def once():
        solver.set_dual_initialization(lam0)
        t = time.perf_counter()
        st = solver.solve(qp_out)
        return time.perf_counter() - t, st, qp_out.iter
    return once


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic"}
    if ref_lib_path(True) is None:
        print(json.dumps(dict(base, unavailable="oracle/_ref/libtreeqp_ref_omp.so missing (built from /root/reference)")))
        return
    prob = workload(args.depth, 0)
    once = run_reference_solve(prob)
    t_full, st, iters_full = once()  # untimed probe = first warm-up step, full solve
    budget = 240.0
    total_steps = args.steps + max(args.warmup - 1, 0)
    if t_full * total_steps <= budget:
        sample, scale, fn = "full solve (%d Newton iterations) per step" % iters_full, 1.0, once
    else:
        n = max(1, int(iters_full * budget / (t_full * total_steps)))
        fn = run_reference_solve(prob, max_iter=n)
        scale = iters_full / float(n)
        sample = "first %d of %d Newton iterations of the same QP per step, time scaled by %d/%d" % (n, iters_full, iters_full, n)
    for _ in range(max(args.warmup - 1, 0)):
        fn()
    times = []
    for _ in range(args.steps):
        t, _, _ = fn()
        times.append(t * scale)
    sec = float(np.sum(times)) / args.steps
    value = 1.0 / sec
    print(json.dumps(dict(
        base, value=value, ms_per_step=1e3 * sec, us_per_newton_iter=1e6 * sec / iters_full,
        config={"workload": "random_qp binary tree depth %d (%d nodes) nx=12 nu=4 |u|<=0.2, TDUNES clipping, 1 tree per GPU" % (args.depth, prob.Nn),
                "newton_iters": iters_full, "status": st, "parallelism": "reference CPU implementation, OpenMP over nodes/blocks, rank 0 only"},
        cpu_baseline={"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        gpu_launches=0)))


# ----------------------------------------------------------------- secondary workloads (N = 1)
def bench_batch(lib, n=4096, reps=3, rank=0, world=1, barrier=None, max_over_ranks=None):
    """config 4: n independent linear_chain tree QPs (175 nodes each) as one forest; with N ranks each
    rank takes n/N instances (QP i -> rank i // (n/N)), no collective on the data path"""
    barrier = barrier or (lambda: None)
    max_over_ranks = max_over_ranks or (lambda x: x)
    nl = n // world
    probs = P.linear_chain_batch(nl, seed=rank * nl)
    qp_ins = [capi.QpIn(lib, p) for p in probs]
    qp_outs = [capi.QpOut(lib, p) for p in probs]
    batch = capi.TdunesBatch(lib, qp_ins)
    assert batch.upload() == capi.TREEQP_OK
    batch.solve_resident()
    t_res = []
    for _ in range(reps):
        barrier(); t0 = time.perf_counter(); batch.solve_resident(); barrier()
        t_res.append(max_over_ranks(time.perf_counter() - t0))
    assert (batch.status == 0).all()
    t_e2e = []
    for _ in range(reps):
        barrier(); t0 = time.perf_counter(); batch.solve(qp_outs); barrier()
        t_e2e.append(max_over_ranks(time.perf_counter() - t0))
    st = batch.stats()
    iters = np.array([o.iter for o in qp_outs])
    n = nl * world
    out = {"workload": "%d x linear_chain nm=4 nu=3 md=3 Nr=2 Nh=20 (175 nodes), |u|<=0.5, x0 seeded%s" % (n, "" if world == 1 else ", %d per GPU" % nl),
           "qp_per_s_resident": n / min(t_res), "qp_per_s_e2e": n / min(t_e2e), "ms_per_batch_resident": 1e3 * min(t_res),
           "ms_per_batch_e2e": 1e3 * min(t_e2e), "newton_iters_max": int(iters.max()), "newton_iters_median": float(np.median(iters)),
           "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"], "launches": st["kernel_launches"]}
    if world > 1:
        out["scaling"] = "strong (4096 QPs in total)"
    path = ref_lib_path(False)
    if path and world == 1:
        ref = capi.TreeQpLib(path)
        m, tt = 32, 0.0
        for i in range(m):
            rin, rout = capi.QpIn(ref, probs[i]), capi.QpOut(ref, probs[i])
            rs = capi.Tdunes(ref, rin)
            rs.set_dual_initialization(np.zeros(probs[i].n_lambda))
            t0 = time.perf_counter(); rs.solve(rout); tt += time.perf_counter() - t0
            assert rout.iter == iters[i]
        out["cpu_reference_qp_per_s_per_core"] = m / tt
        out["cpu_sample"] = "first %d QPs of the batch, 1 thread, unmodified reference; iteration counts identical" % m
    batch.destroy()
    return out


def bench_sdunes(lib, reps=10):
    """config 2: spring-mass SDUNES, md=3 Nr=4 Nh=20 -> 81 scenarios, 1417 nodes"""
    prob = P.eliminate_x0(P.spring_mass(3, 4, 20))
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    s = capi.Sdunes(lib, qp_in)
    z = (np.zeros(s.n_lambda), np.zeros(s.n_mu))
    s.set_dual_initialization(*z)
    s.upload()
    best, e2e = 1e9, 1e9
    for _ in range(reps):
        s.solve_resident(); best = min(best, s.stats()["gpu_ms"])
    for _ in range(reps):
        s.set_dual_initialization(*z)
        t0 = time.perf_counter(); st = s.solve(qp_out); e2e = min(e2e, 1e3 * (time.perf_counter() - t0))
    stats = s.stats()
    out = {"workload": "spring-mass SDUNES md=3 Nr=4 Nh=20 (81 scenarios, 1417 nodes)", "status": st, "newton_iters": qp_out.iter,
           "ms_per_solve_resident": best, "ms_per_solve_e2e": e2e, "solves_per_s_resident": 1e3 / best,
           "us_per_newton_iter": 1e3 * best / max(qp_out.iter, 1), "kernel_launches": stats["kernel_launches"],
           "phases_ms": {k: round(v[0], 4) for k, v in stats["kernels"].items()}}
    path = ref_lib_path(False)
    if path:
        ref = capi.TreeQpLib(path)
        rin, rout = capi.QpIn(ref, prob), capi.QpOut(ref, prob)
        rs = capi.Sdunes(ref, rin)
        tb = 1e9
        for _ in range(reps):
            rs.set_dual_initialization(*z); rs.solve(rout); tb = min(tb, 1e3 * rout.total_time)
        out["cpu_reference_ms_per_solve"] = tb
        out["cpu_reference_newton_iters"] = rout.iter
    s.destroy()
    return out


def bench_dense(lib, reps=5):
    """config 5: dense stage QPs (qp_solver = TREEQP_QPOASES_SOLVER): linear_chain nm=8 nu=7 (nx=16) md=4
    Nr=2 Nh=20 with dense Q/R/S blocks, plus trees of the quadcopter (nx=6, nu=4) and crane (nx=4, nu=1)
    dimensions, md=3 Nr=2 Nh=20"""
    def ubox(p, ub):
        p.umin[:] = -ub; p.umax[:] = ub; p.xmin[:] = -1e12; p.xmax[:] = 1e12
        return p
    probs = {
        "linear_chain_nm8_nu7_md4_Nh20": P.densify(ubox(P.linear_chain(8, 7, 4, 2, 20, seed=2), 0.3), 6, 0.1, 0.0),
        "quadcopter_dims_nx6_nu4_md3_Nh20": P.densify(P.random_tree(P.multistage_nk(3, 2, 20), 6, 4, seed=8, ubound=0.3), 4, 0.1, 0.0),
        "crane_dims_nx4_nu1_md3_Nh20": P.densify(P.random_tree(P.multistage_nk(3, 2, 20), 4, 1, seed=9, ubound=0.2), 5, 0.05, 0.0),
    }
    out = {}
    path = ref_lib_path(False)
    ref = capi.TreeQpLib(path) if path else None
    for name, prob in probs.items():
        qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
        s = capi.Tdunes(lib, qp_in, qp_solver=1, maxIter=200)
        lam0 = np.zeros(prob.n_lambda)
        s.set_dual_initialization(lam0)
        assert s.upload() == capi.TREEQP_OK
        best, e2e = 1e9, 1e9
        for _ in range(reps):
            st = s.solve_resident(); best = min(best, s.stats()["gpu_ms"])
        for _ in range(reps):
            s.set_dual_initialization(lam0)
            t0 = time.perf_counter(); st = s.solve(qp_out); e2e = min(e2e, 1e3 * (time.perf_counter() - t0))
        rec = {"nodes": int(prob.Nn), "status": int(st), "newton_iters": int(qp_out.iter), "ms_per_solve_resident": best,
               "ms_per_solve_e2e": e2e, "us_per_newton_iter": 1e3 * best / max(qp_out.iter, 1),
               "kkt": float(lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))}
        s.destroy()
        if ref is not None:
            rin, rout = capi.QpIn(ref, prob), capi.QpOut(ref, prob)
            rs = capi.Tdunes(ref, rin, qp_solver=1, maxIter=200)
            tb = 1e9
            for _ in range(3):
                rs.set_dual_initialization(lam0); rs.solve(rout); tb = min(tb, 1e3 * rout.total_time)
            rec["cpu_reference_ms_per_solve"] = tb
            rec["cpu_reference_newton_iters"] = int(rout.iter)
        out[name] = rec
    # the GPU regime of small dense trees is the batch: 256 QPs of the quadcopter dimensions as one forest
    try:
        n = 256
        bp = [P.densify(P.random_tree(P.multistage_nk(3, 2, 20), 6, 4, seed=100 + i, ubound=0.3), 4, 0.1, 0.0) for i in range(n)]
        qi = [capi.QpIn(lib, p) for p in bp]
        qo = [capi.QpOut(lib, p) for p in bp]
        b = capi.TdunesBatch(lib, qi, qp_solver=1, maxIter=200)
        assert b.upload() == capi.TREEQP_OK
        b.solve_resident()
        t_res = []
        for _ in range(3):
            t0 = time.perf_counter(); b.solve_resident(); t_res.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); b.solve(qo); t_e2e = time.perf_counter() - t0
        iters = np.array([o.iter for o in qo])
        rec = {"workload": "%d x dense tree nx=6 nu=4 md=3 Nr=2 Nh=20 (175 nodes) as one forest" % n, "converged": int((b.status == 0).sum()),
               "qp_per_s_resident": n / min(t_res), "qp_per_s_e2e": n / t_e2e, "ms_per_batch_resident": 1e3 * min(t_res),
               "newton_iters_median": float(np.median(iters)), "newton_iters_max": int(iters.max())}
        if ref is not None:
            m, tt = 8, 0.0
            for i in range(m):
                rin, rout = capi.QpIn(ref, bp[i]), capi.QpOut(ref, bp[i])
                rs = capi.Tdunes(ref, rin, qp_solver=1, maxIter=200)
                rs.set_dual_initialization(np.zeros(bp[i].n_lambda)); rs.solve(rout); tt += rout.total_time
                rec.setdefault("cpu_iters_match", True)
                rec["cpu_iters_match"] = bool(rec["cpu_iters_match"] and rout.iter == iters[i])
            rec["cpu_reference_qp_per_s_per_core"] = m / tt
        b.destroy()
        out["batch256_quadcopter_dims"] = rec
    except Exception as e:
        out["batch256_quadcopter_dims"] = {"error": repr(e)}
    return out


def bench_mpc(lib, steps=20):
    """closed-loop re-solves.  Between consecutive QPs only x0 changes
    (benchmark/treeqp_main.m:235-237): set_x0 -> update_x0 (root + children travel) -> dual start =
    previous multipliers (device to device) -> solve -> download, against a full treeqp_tdunes_solve
    per sample and against the reference doing the same loop on one host core."""
    prob = P.spring_mass(3, 4, 20)                      # 1417 nodes, x0 pinned on the root
    dims = P.eliminate_x0(prob)
    rng = np.random.default_rng(5)
    x0s = [np.array([0.01, 0.02, 0.03, 0.04]) + 0.002 * rng.standard_normal(4) for _ in range(steps + 1)]
    qi, qo = capi.QpIn(lib, prob), capi.QpOut(lib, dims)
    qi.eliminate_x0()
    s = capi.Tdunes(lib, qi)
    s.set_dual_initialization(np.zeros(dims.n_lambda))
    assert s.upload() == capi.TREEQP_OK and s.solve_resident() == 0
    iters, t0 = [], time.perf_counter()
    for k in range(steps):
        qi.set_x0(x0s[k + 1])
        assert s.update_x0() == capi.TREEQP_OK and s.warm_start_resident() == capi.TREEQP_OK
        assert s.solve_resident() == 0 and s.download(qo) == capi.TREEQP_OK
        iters.append(s.stats()["iter"])
    t_fast = (time.perf_counter() - t0) / steps
    h2d_fast = s.stats()["h2d_bytes"]
    lam = qo.lam.copy()
    t0 = time.perf_counter()
    for k in range(steps):
        qi.set_x0(x0s[k + 1])
        s.set_dual_initialization(lam)
        assert s.solve(qo) == 0
        lam = qo.lam.copy()
    t_full = (time.perf_counter() - t0) / steps
    out = {"workload": "spring-mass md=3 Nr=4 Nh=20 (1417 nodes), %d closed-loop samples, x0 perturbed, warm-started duals" % steps,
           "ms_per_sample_update_x0_path": 1e3 * t_fast, "ms_per_sample_full_solve_call": 1e3 * t_full,
           "h2d_bytes_update_x0": h2d_fast, "h2d_bytes_full": s.stats()["h2d_bytes"], "newton_iters_median": float(np.median(iters))}
    s.destroy()
    path = ref_lib_path(False)
    if path:
        ref = capi.TreeQpLib(path)
        ri, ro = capi.QpIn(ref, prob), capi.QpOut(ref, dims)
        ri.eliminate_x0()
        rs = capi.Tdunes(ref, ri)
        rs.set_dual_initialization(np.zeros(dims.n_lambda)); rs.solve(ro)
        lam, rit, t0 = ro.lam.copy(), [], time.perf_counter()
        for k in range(steps):
            ri.set_x0(x0s[k + 1])
            rs.set_dual_initialization(lam)
            assert rs.solve(ro) == 0
            lam = ro.lam.copy(); rit.append(ro.iter)
        out["cpu_reference_ms_per_sample"] = 1e3 * (time.perf_counter() - t0) / steps
        out["cpu_reference_iters_median"] = float(np.median(rit))
    return out


def bench_sharded(lib, args, rank, world, local, dist, barrier, max_over_ranks):
    """N > 1: ONE tree (the N=1 workload) sharded by subtree over the ranks: strong scaling"""
    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    comm = capi.Comm(lib, rank, world, local, bcast)
    prob = workload(args.depth, 0)
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    sh = capi.TdunesSharded(lib, qp_in, comm, **OPTS)
    sh.set_dual_initialization(np.zeros(prob.n_lambda))
    assert sh.upload() == capi.TREEQP_OK
    for _ in range(max(args.warmup, 1)):
        st = sh.solve_resident()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = sh.solve_resident()
    barrier()
    t = max_over_ranks(time.perf_counter() - t0)
    stats, plan = sh.stats(), sh.plan()
    lib.lib.treeqp_b200_set_profiling(1)
    sh.solve_resident()
    prof = sh.stats()
    lib.lib.treeqp_b200_set_profiling(0)
    comm_ms = prof["kernels"].get("nccl+pack", (0.0, 0))
    out = {"workload": "the N=1 tree (%d nodes) sharded by subtree over %d GPUs" % (prob.Nn, world), "scaling": "strong",
           "status": st, "newton_iters": stats["iter"], "ls_trials": stats["ls_total"], "ms_per_solve": 1e3 * t / args.steps,
           "solves_per_s": args.steps / t, "us_per_newton_iter": 1e6 * t / args.steps / max(stats["iter"], 1),
           "cut_depth": plan["cut_depth"], "cut_nodes": plan["ncut"], "local_nodes": plan["n_local"],
           "collectives_ms_per_solve": comm_ms[0], "collectives_per_solve": comm_ms[1]}
    sh.destroy()
    comm.destroy()
    return out


# ------------------------------------------------------------------------------------- own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--depth", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (config 4 batch, config 2 SDUNES)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # host threads of the packer (OpenMP): torchrun exports OMP_NUM_THREADS=1 to every rank, which
    # would serialise the BLASFEO -> SoA conversion of the e2e path; give each rank its share of cores
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)

    lib = capi.TreeQpLib(capi.PRODUCT_LIB)
    assert lib.lib.treeqp_b200_set_device(local) == 0
    # every rank solves its own copy of the SAME QP (seed 0): per-GPU work is then identical at every
    # N (the Newton iteration count depends on the data), which is what weak scaling compares
    prob = workload(args.depth, 0)
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    solver = capi.Tdunes(lib, qp_in, **OPTS)
    lam0 = np.zeros(prob.n_lambda)
    solver.set_dual_initialization(lam0)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident (inputs already in HBM): value
    assert solver.upload() == capi.TREEQP_OK
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        st = solver.solve_resident()
    barrier()
    clk_lo = clocks.mark()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        st = solver.solve_resident()
        dev_ms += solver.stats()["gpu_ms"]
    barrier()
    t_res = max_over_ranks(time.perf_counter() - t0)
    stats = solver.stats()
    assert st == 0, "benchmark QP did not converge: status %d" % st
    iters = stats["iter"]
    launches = stats["kernel_launches"] * args.steps

    # ---- end to end through the drop-in call with host buffers
    for _ in range(min(args.warmup, 1)):
        solver.set_dual_initialization(lam0)
        solver.solve(qp_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        solver.set_dual_initialization(lam0)
        st2 = solver.solve(qp_out)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_stats = solver.stats()
    clk = clocks.stop(clk_lo, clocks.mark()) if rank == 0 else None
    assert st2 == 0 and qp_out.iter == iters
    kkt = float(lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
    kkt_dev = solver.kkt_resident()  # the same certificate evaluated on the device (k_kkt)

    # ---- roofline of the dominant kernel class: one profiled solve (CUDA events per launch)
    lib.lib.treeqp_b200_set_profiling(1)
    solver.solve_resident()
    prof = solver.stats()
    lib.lib.treeqp_b200_set_profiling(0)

    sharded = None
    if dist is not None:
        try:
            sharded = bench_sharded(lib, args, rank, world, local, dist, barrier, max_over_ranks)
        except Exception as e:  # report, never hide
            sharded = {"error": repr(e)}

    batch_sharded = None
    if dist is not None and not args.no_extra:
        try:
            batch_sharded = bench_batch(lib, 4096, 3, rank, world, barrier, max_over_ranks)
        except Exception as e:
            batch_sharded = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks, peak_src = float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    alg = algorithmic_bytes(prob)
    lpc = launches_per_call(prob)
    kern = prof["kernels"]
    fused = "k_hessian" not in kern
    if fused:
        # uniform trees run k_build_factor: dual gradient + Hessian blocks + Cholesky + Schur complement
        # in one kernel per depth (event class "k_factor_level"), i.e. SURVEY 8d's K2 + K3 together
        alg["k_factor_level"] = alg["k_gradient"] + alg["k_hessian"] + alg["k_factor_level"]
    top = max((k for k in kern if k in alg), key=lambda k: kern[k][0])
    t_ms, n_launch = kern[top]
    calls = n_launch / lpc[top]
    achieved = alg[top] * calls / (t_ms * 1e-3) / 1e9
    breakdown = {k: {"ms": round(v[0], 4), "launches": v[1],
                     "GBps": round(alg[k] * (v[1] / lpc[k]) / (v[0] * 1e-3) / 1e9, 1) if k in alg else None}
                 for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
        if tr.get("depth") == args.depth and top in tr["per_launch_bytes"]:
            traffic, traffic_src = tr["per_launch_bytes"][top], tr["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": ("k_build_factor (fused K2+K3, one launch per depth)" if fused and top == "k_factor_level" else top),
                "achieved": achieved, "peak": peaks, "unit": "GB/s",
                "frac": achieved / peaks, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg[top] / lpc[top], "peak_source": peak_src,
                "algorithmic_bytes_per_call": alg[top], "launches_per_call": lpc[top],
                "avg_launch_us": 1e3 * t_ms / n_launch, "kernel_share_of_solve": t_ms / prof["gpu_ms"],
                "per_kernel": breakdown}

    cpu_baseline = None
    if not args.no_cpu_baseline and args.gpus == 1 and ref_lib_path(True):
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        once = run_reference_solve(prob)
        t_cpu, st_cpu, it_cpu = once()
        cpu_baseline = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": cores, "kind": "reference",
                        "sample": "1 full solve of the same QP (%d Newton iterations, status %d), OpenMP build of the unmodified reference" % (it_cpu, st_cpu),
                        "us_per_newton_iter": 1e6 * t_cpu / max(it_cpu, 1), "newton_iters": it_cpu}
    elif not args.no_cpu_baseline and args.gpus == 1:
        cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}

    extra = {}
    if batch_sharded is not None:
        extra["batch4096_linear_chain"] = batch_sharded
    if world == 1 and not args.no_extra:
        for name, fn in (("batch4096_linear_chain", bench_batch), ("sdunes_spring_mass_cfg2", bench_sdunes),
                         ("dense_stage_qps_cfg5", bench_dense), ("mpc_resolve_x0_fast_path", bench_mpc)):
            try:
                extra[name] = fn(lib)
            except Exception as e:
                extra[name] = {"error": repr(e)}

    n = world
    out = {
        "metric": METRIC, "value": n * args.steps / t_res, "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "random_qp binary tree depth %d (%d nodes) nx=12 nu=4 |u|<=0.2, TDUNES clipping, 1 tree per GPU" % (args.depth, prob.Nn),
                   "newton_iters": iters, "ls_trials": stats["ls_total"], "kkt": kkt, "kkt_device": kkt_dev,
                   "l2": "inputs %.2f GB + iterates/Hessian blocks, all >> 126 MB L2: no flush between steps" % (stats["h2d_bytes"] / 1e9),
                   "parallelism": "1 QP per GPU, no collective"},
        "us_per_newton_iter": 1e3 * (1e3 * t_res / args.steps) / max(iters, 1),
        "device_ms_per_step": dev_ms / args.steps,
        "e2e": {"value": n * args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": e2e_stats["h2d_bytes"],
                "d2h_bytes_per_step": e2e_stats["d2h_bytes"], "ms_per_step": 1e3 * t_e2e / args.steps,
                "pack_ms": e2e_stats["pack_ms"], "h2d_ms": e2e_stats["h2d_ms"], "d2h_ms": e2e_stats["d2h_ms"],
                "gpu_ms": e2e_stats["gpu_ms"]},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "other_workloads": extra, "sharded_single_tree": sharded,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

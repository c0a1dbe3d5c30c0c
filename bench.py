#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native TDUNES hot path.

Metric (BASELINE.json): tree-QP solves/sec (and us per dual Newton iteration) on the 10^5-node
random tree (config 3: full binary tree of depth 16 = 131071 nodes, nx=12, nu=4, clipping stage
QPs, FP64).  One "step" = one complete dual Newton solve of that QP from the dual start lambda=0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--depth D]

  value  whole-job throughput with the QP data already resident in HBM
         (N = 1: treeqp_tdunes_solve_resident; N > 1: the SAME tree sharded by subtree over the N GPUs,
         treeqp_tdunes_sharded_solve_resident -- strong scaling, BASELINE config 3)
  e2e    same metric through the call with HOST buffers (treeqp_tdunes_solve / ..._sharded_solve:
         BLASFEO views -> pinned SoA -> H2D -> Newton loop -> D2H -> BLASFEO views), every step
  parity the solution of the timed QP against the unmodified reference's solution of the same QP (status,
         Newton iterations, active sets, x/u/lam/mu to 1e-10 relative); the run FAILS if it differs
  roofline      dominant kernel class: SURVEY.md 8d bytes (K2 + K3 for the fused build/factor kernel) over
                the CUDA-event time of its launches; frac_dram = measured DRAM bytes (ncu) over the same
                time; compute = factor/solve FLOPs over the measured FP64 peak
  cpu_baseline  the unmodified reference (oracle/_ref, OpenMP build, all host cores) on the same QP,
                min over 5 solves

--impl reference: times the reference's own CPU implementation (rank 0 only, all host cores).
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
    """bytes each phase must move per call, from the problem dimensions only: the K1..K5 formulas of
    SURVEY.md 8(d) verbatim (doubles x 8), plus the vector streams of one line-search pass and the
    factor + substitution FLOPs of the same section"""
    nx, nu, nk = prob.nx.astype(np.int64), prob.nu.astype(np.int64), prob.nk.astype(np.int64)
    dad = prob.dad
    Nn = prob.Nn
    ab = np.zeros(Nn, dtype=np.int64)
    ab[1:] = nx[1:] * (nx[dad[1:]] + nu[dad[1:]])
    kid_ab = np.bincount(dad[1:], weights=ab[1:], minlength=Nn).astype(np.int64)
    d = np.bincount(dad[1:], weights=nx[1:], minlength=Nn).astype(np.int64)
    par = nk > 0
    nrp = par.copy(); nrp[0] = False            # non-root parents
    k1 = 8 * int(((5 * nx + 4 * nu) + kid_ab + d + 4 * nx + 4 * nu).sum())
    k2 = 8 * int((ab[1:] + 4 * nx[1:]).sum() + (2 * nx + 2 * nu + d * (d + 1) // 2)[par].sum() + (nx * d)[nrp].sum())
    k3 = 8 * int((d * (d + 1) + 2 * d)[par].sum() + (2 * nx * d + nx * (nx + 1) + 2 * nx)[nrp].sum())
    k4 = 8 * int((d * (d + 1) // 2 + 2 * d)[par].sum() + (nx * d + nx)[nrp].sum())
    k5 = 8 * int(((6 * nx + 5 * nu) + kid_ab + 2 * d + nx + nu).sum())
    model = 8 * int((2 * nx + kid_ab + 2 * d + nx + nu + 2).sum())      # k_stage<2>: affine direction model
    trial = 8 * int(6 * nx.sum() + 6 * nu.sum() + 2 * Nn)               # one k_ls_eval pass (vectors only)
    flops = int((d ** 3 // 3 + 3 * d * d)[par].sum() + (nx * d * d + nx * nx * d + 2 * nx * d)[nrp].sum())
    return {"K1": k1, "K2": k2, "K3": k3, "K4": k4, "K5": k5, "model": model, "trial": trial, "factor_flops": flops,
            "iteration_L1": k1 + k2 + k3 + k4 + 2 * k5}


def class_bytes(alg, fused):
    """algorithmic bytes per CALL of each event class (names of device.cu)"""
    if fused:
        return {"k_stage_fused<0>": alg["K1"], "k_stage_fused<2>": alg["model"], "k_ls_eval": alg["trial"],
                "k_build_factor": alg["K2"] + alg["K3"], "k_backsolve_fused": alg["K4"]}
    return {"k_stage<0>": alg["K1"], "k_stage<2>": alg["model"], "k_trial_dense+k_ls_eval": alg["trial"],
            "k_gradient+k_hessian": alg["K2"], "k_factor_level": alg["K3"], "k_backsolve_level": alg["K4"]}


def launches_per_call(prob, kern):
    """how many launches make up ONE call of each phase (level kernels: one per depth)"""
    stage = P.tree_tables(prob.nk)[1]
    Nh = int(stage.max())
    out = {k: 1 for k in kern}
    for k in ("k_build_factor", "k_factor_level"):
        out[k] = Nh
    for k in ("k_backsolve_fused", "k_backsolve_level"):
        out[k] = max(Nh - 1, 1)
    return out


def measured_peaks():
    """HBM GB/s from the driver's MEASURED_PEAKS.json; FP64 TFLOP/s from this repo's own microbenchmark
    (tools/micro/dmma_rate.cu on a B200 of this pool, profiles/r02_micro_fp64_peak.txt: DFMA 33.9, DMMA 37.1)"""
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm, src = float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return hbm, src, 33.9, "measured DFMA rate, tools/micro/dmma_rate.cu (profiles/r02_micro_fp64_peak.txt)"


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def parity_block(prob, sol, ref, against):
    """sol / ref: dicts with status, iter, x, u, lam, mu_x, mu_u (flat).  north_star bar: identical status,
    Newton iteration count and active sets; primal/dual solution within 1e-10 relative"""
    errs = {k: rel_err(sol[k], ref[k]) for k in ("x", "u", "lam", "mu_x", "mu_u")}
    act = True
    n_active = 0
    for k, lo, hi in (("x", prob.xmin, prob.xmax), ("u", prob.umin, prob.umax)):
        act = act and bool(np.array_equal(sol[k] == lo, ref[k] == lo)) and bool(np.array_equal(sol[k] == hi, ref[k] == hi))
        n_active += int((ref[k] == lo).sum() + (ref[k] == hi).sum())
    ok = (int(sol["status"]) == int(ref["status"]) and int(sol["iter"]) == int(ref["iter"]) and act
          and all(e < 1e-10 for e in errs.values()))
    return {"ok": bool(ok), "against": against, "tolerance": 1e-10, "status": [int(sol["status"]), int(ref["status"])],
            "newton_iters": [int(sol["iter"]), int(ref["iter"])], "active_sets_identical": bool(act), "active_bounds": n_active,
            "max_rel_err": {k: float("%.3e" % v) for k, v in errs.items()}}


def out_dict(qp_out, status):
    return {"status": int(status), "iter": int(qp_out.iter), "x": qp_out.x, "u": qp_out.u, "lam": qp_out.lam,
            "mu_x": qp_out.mu_x, "mu_u": qp_out.mu_u}


def omp_threads(n=None):
    """set (n given) and read the thread count the OpenMP runtime will really use"""
    import ctypes
    try:
        g = ctypes.CDLL("libgomp.so.1")
        if n:
            g.omp_set_num_threads(int(n))
        return int(g.omp_get_max_threads())
    except Exception:
        return None


def config_dict(depth, Nn, n_gpus):
    """identical in both arms: the workload BASELINE.json's metric is quoted on"""
    return {"workload": "random_qp binary tree depth %d (%d nodes) nx=12 nu=4 |u|<=0.2, TDUNES clipping, FP64, dual start 0, maxIter 200" % (depth, Nn),
            "sharding": "one GPU" if n_gpus == 1 else "the one tree sharded by subtree over %d GPUs" % n_gpus,
            "l2": "inputs 0.30 GB + iterates/factors per solve, all >> 126 MB L2: no flush between steps"}


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
    """treeqp_tdunes_solve of the UNMODIFIED reference (OpenMP build, all host cores): returns a closure
    that solves once from lambda = 0 and the qp_out it fills"""
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
    once.qp_out = qp_out
    once.keep = (lib, qp_in, solver)
    return once


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the reference gets ALL host cores, set before
    # its library (and with it the OpenMP runtime) is loaded, and checked against the runtime afterwards
    os.environ["OMP_NUM_THREADS"] = str(cores)
    prob = workload(args.depth, 0)
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args.depth, prob.Nn, args.gpus)}
    if ref_lib_path(True) is None:
        print(json.dumps(dict(base, unavailable="oracle/_ref/libtreeqp_ref_omp.so missing (built from /root/reference)")))
        return
    once = run_reference_solve(prob)
    threads = omp_threads(cores)
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
        detail={"newton_iters": iters_full, "status": st, "best_ms": 1e3 * float(np.min(times)),
                "parallelism": "reference CPU implementation, OpenMP over nodes/blocks, rank 0 only, %s OpenMP threads" % threads},
        cpu_baseline={"value": value, "unit": UNIT, "cores": threads or cores, "host_cores": cores, "kind": "reference", "sample": sample},
        e2e={"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        gpu_launches=0)))


def solve_roofline(prob, newton_iters_total, seconds, io_bytes, ls_per_iter=1.0):
    """whole-solve HBM roofline of a secondary workload: SURVEY 8(d) bytes of one Newton iteration
    (K1+K2+K3+K4+(L+1)K5, L = line-search trials per iteration) x iterations over the resident time,
    next to the compulsory traffic (inputs + outputs once per solve) for the on-chip regime"""
    hbm, hbm_src, _, _ = measured_peaks()
    alg = algorithmic_bytes(prob)
    per_it = alg["K1"] + alg["K2"] + alg["K3"] + alg["K4"] + (ls_per_iter + 1.0) * alg["K5"]
    ach = per_it * newton_iters_total / seconds / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": None,
            "bytes_per_newton_iter_per_qp": per_it, "newton_iters_total": newton_iters_total,
            "compulsory_io_GBps": io_bytes / seconds / 1e9, "peak_source": hbm_src,
            "note": "small trees are latency/on-chip bound: SURVEY 8(d) asks for the HBM line for batch throughput"}


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
           "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"], "launches": st["kernel_launches"],
           # per GPU: this rank's QPs over the (max over ranks) time, against ONE GPU's HBM peak
           "roofline": solve_roofline(probs[0], int(iters.sum()), min(t_res), st["h2d_bytes"] + st["d2h_bytes"],
                                      st["ls_total"] / max(int(iters.sum()), 1))}
    if world == 1:
        try:
            out["mpc_step_x0_only"] = bench_batch_x0(lib, min(n, 1024))
        except Exception as e:
            out["mpc_step_x0_only"] = {"error": repr(e)}
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


def bench_batch_x0(lib, n):
    """closed-loop step of a batch of MPC QPs: between two samples only x0 of every QP changes
    (benchmark/treeqp_main.m:235-237); treeqp_tdunes_batch_update_x0 re-sends b of the root's children instead
    of the whole problem.  Timed: update_x0 + solve_resident + download (set_x0 on the host containers is the
    caller's data preparation, like building qp_in for the other e2e numbers)."""
    rng = np.random.default_rng(11)
    probs = [P.linear_chain(4, 3, 3, 2, 20, seed=i, eliminate_x0=False) for i in range(n)]
    for p in probs:
        p.umin[:] = -0.5; p.umax[:] = 0.5
        nx0 = int(p.nx[0])
        p.xmin[nx0:] = -1e12; p.xmax[nx0:] = 1e12
    qis = [capi.QpIn(lib, p) for p in probs]
    for q in qis:
        q.eliminate_x0()
    dims = [P.eliminate_x0(p) for p in probs]
    qos = [capi.QpOut(lib, d) for d in dims]
    b = capi.TdunesBatch(lib, qis)
    assert b.upload() == capi.TREEQP_OK and b.solve_resident() == capi.TREEQP_OK
    full = b.stats()["h2d_bytes"]
    ts = []
    for step in range(3):
        for q, p in zip(qis, probs):
            q.set_x0(p.xmin[:int(p.nx[0])] * (0.9 + 0.1 * rng.random()))
        t0 = time.perf_counter()
        assert b.update_x0() == capi.TREEQP_OK
        h2d = b.stats()["h2d_bytes"]
        assert b.solve_resident() == capi.TREEQP_OK and b.download(qos) == capi.TREEQP_OK
        ts.append(time.perf_counter() - t0)
    out = {"workload": "%d x linear_chain (175 nodes), x0 of every QP changed, update_x0 + solve_resident + download" % n,
           "qp_per_s": n / min(ts), "ms_per_step": 1e3 * min(ts), "h2d_bytes_update_x0": h2d, "h2d_bytes_full_upload": full,
           "converged": int((b.status == 0).sum())}
    b.destroy()
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
           "phases_ms": {k: round(v[0], 4) for k, v in stats["kernels"].items()},
           "roofline": solve_roofline(prob, int(qp_out.iter), 1e-3 * best, stats["h2d_bytes"] + stats["d2h_bytes"],
                                      stats["ls_total"] / max(int(qp_out.iter), 1))}
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
    # the quadcopter and crane MODELS of the benchmark harness (problems.quadcopter / problems.crane restate
    # benchmark/quadcopter/initialize_quadcopter.m:33-114 and benchmark/crane/initialize_crane.m:33-72), the
    # instances of the parity goldens (tests/cases.py: quadcopter_cfg5, crane_cfg5), dense stage Hessians
    quad_x0, crane_x0 = [0.03, -0.03, 0.01, 0.8, -0.8, 0.3], [0.0, 0.2, 0.1, -0.3]
    probs = {
        "linear_chain_nm8_nu7_md4_Nh20": P.densify(ubox(P.linear_chain(8, 7, 4, 2, 20, seed=2), 0.3), 6, 0.1, 0.0),
        "quadcopter_md3_Nr2_Nh20": P.densify_scaled(P.quadcopter(3, 2, 20, i_mpc=14, x0=quad_x0), 4),
        "crane_md3_Nr2_Nh20": P.densify_scaled(P.crane(3, 2, 20, x0=crane_x0), 5),
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
        lib.lib.treeqp_b200_set_profiling(1)
        s.solve_resident()
        rec["profiled_kernel_ms"] = {k: (round(v[0], 3), v[1]) for k, v in sorted(s.stats()["kernels"].items(), key=lambda kv: -kv[1][0])}
        rec["launches_per_solve"] = s.stats()["kernel_launches"]
        lib.lib.treeqp_b200_set_profiling(0)
        sst = s.stats()
        rec["roofline"] = solve_roofline(prob, int(qp_out.iter), 1e-3 * best, sst["h2d_bytes"] + sst["d2h_bytes"],
                                         sst["ls_total"] / max(int(qp_out.iter), 1))
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
        bp = []
        for i in range(n):   # closed-loop states of different samples: x0 = quad_x0 .* U(0.2, 0.6), seeded
            x0 = np.array(quad_x0) * np.random.Generator(np.random.PCG64(100 + i)).uniform(0.2, 0.6, 6)
            bp.append(P.densify_scaled(P.quadcopter(3, 2, 20, i_mpc=14, x0=x0), 4))
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
        rec = {"workload": "%d x quadcopter tree (dense stage QPs, nx=6 nu=4 md=3 Nr=2 Nh=20, 175 nodes) as one forest, x0 seeded" % n, "converged": int((b.status == 0).sum()),
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
                rec.setdefault("cpu_converged_of_8", 0)
                rec["cpu_converged_of_8"] += int(rout.iter < 200)
            rec["cpu_reference_qp_per_s_per_core"] = m / tt
        b.destroy()
        out["batch256_quadcopter"] = rec
    except Exception as e:
        out["batch256_quadcopter"] = {"error": repr(e)}
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


def run_single(lib, args, prob, barrier, max_over_ranks, clocks, rank):
    """N = 1 headline (and, at N > 1, the replica line: every rank solves its own copy of the tree)"""
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    # the caller's tree_qp_in / tree_qp_out buffers are page-locked once (treeqp_b200_host_register, the one line
    # INTEGRATION.md asks of an MPC loop): e2e copies then run by DMA straight from / into them
    pinned = bool(qp_in.pin() and qp_out.pin()) if os.environ.get("TREEQP_B200_BENCH_UNPINNED", "0") != "1" else False
    solver = capi.Tdunes(lib, qp_in, **OPTS)
    lam0 = np.zeros(prob.n_lambda)
    solver.set_dual_initialization(lam0)
    # ---- resident (inputs already in HBM): value
    assert solver.upload() == capi.TREEQP_OK
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
    clk_hi = clocks.mark()
    assert st2 == 0 and qp_out.iter == stats["iter"]
    kkt = float(lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
    kkt_dev = solver.kkt_resident()  # the same certificate evaluated on the device (k_kkt)
    # ---- profiled solves for the roofline block: level 2 = CUDA events around every launch, but ONE pair
    # around each PDL-chained level sweep (the sweep is timed the way it runs); level 1 = a pair around
    # every launch (serialises the level kernels), kept next to it
    lib.lib.treeqp_b200_set_profiling(2)
    solver.solve_resident()
    prof = solver.stats()
    lib.lib.treeqp_b200_set_profiling(1)
    solver.solve_resident()
    prof["serialised"] = solver.stats()["kernels"]
    lib.lib.treeqp_b200_set_profiling(0)
    sol = out_dict(qp_out, st2)
    # the same call with the buffers NOT page-locked (every byte through the host packer and the library's own pinned
    # staging block): reported next to the headline e2e
    t_unp = None
    if pinned:
        qp_in.mem.unpin(); qp_out.mem.unpin()
        solver.set_dual_initialization(lam0); solver.solve(qp_out)
        t0 = time.perf_counter()
        for _ in range(max(2, args.steps // 4)):
            solver.set_dual_initialization(lam0)
            solver.solve(qp_out)
        t_unp = (time.perf_counter() - t0) / max(2, args.steps // 4)
    solver.destroy()
    return dict(t_res=t_res, t_e2e=t_e2e, dev_ms=dev_ms, stats=stats, e2e_stats=e2e_stats, prof=prof, sol=sol, pinned=pinned, t_e2e_unpinned=t_unp,
                kkt=kkt, kkt_dev=kkt_dev, clk=(clk_lo, clk_hi), launches=stats["kernel_launches"] * args.steps)


def run_sharded(lib, args, prob, rank, world, local, dist, barrier, max_over_ranks, clocks):
    """N > 1 headline: ONE tree (the N=1 workload) sharded by subtree over the ranks: strong scaling"""
    import torch

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    comm = capi.Comm(lib, rank, world, local, bcast)
    qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
    # as at N = 1: the caller's buffers are page-locked once, every rank moves its subtrees by DMA from / into them
    pinned = bool(qp_in.pin() and qp_out.pin()) if os.environ.get("TREEQP_B200_BENCH_UNPINNED", "0") != "1" else False
    sh = capi.TdunesSharded(lib, qp_in, comm, **OPTS)
    lam0 = np.zeros(prob.n_lambda)
    sh.set_dual_initialization(lam0)
    assert sh.upload() == capi.TREEQP_OK
    for _ in range(max(args.warmup, 1)):
        st = sh.solve_resident()
    barrier()
    clk_lo = clocks.mark()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        st = sh.solve_resident()
        dev_ms += sh.stats()["gpu_ms"]
    barrier()
    t_res = max_over_ranks(time.perf_counter() - t0)
    stats, plan = sh.stats(), sh.plan()
    assert st == 0, "benchmark QP did not converge: status %d" % st
    # ---- end to end: every rank packs + uploads its subtrees, solves, downloads its part
    sh.set_dual_initialization(lam0)
    sh.solve(qp_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sh.set_dual_initialization(lam0)
        st2 = sh.solve(qp_out)
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_stats = sh.stats()
    clk_hi = clocks.mark()
    assert st2 == 0 and qp_out.iter == stats["iter"]
    # ---- assemble the solution on every rank: owned parts are disjoint, a SUM of masked vectors is the solution
    own = capi.shard_plan(lib, qp_in, rank, world)["owned"]
    nxm, num, lamm = np.repeat(own, prob.nx), np.repeat(own, prob.nu), np.repeat(own[1:], prob.nx[1:])
    sol = {"status": int(st2), "iter": int(qp_out.iter)}
    for key, mask in (("x", nxm), ("u", num), ("lam", lamm), ("mu_x", nxm), ("mu_u", num)):
        t = torch.from_numpy(np.where(mask, getattr(qp_out, key), 0.0)).cuda()
        dist.all_reduce(t)
        sol[key] = t.cpu().numpy()
    cnt = torch.tensor(own.astype(np.int64)).cuda()
    dist.all_reduce(cnt)
    sol["owners_per_node_ok"] = bool((cnt.cpu().numpy() == 1).all())
    h2d = torch.tensor([float(e2e_stats["h2d_bytes"]), float(e2e_stats["d2h_bytes"])], dtype=torch.float64, device="cuda")
    dist.all_reduce(h2d)
    e2e_stats = dict(e2e_stats, h2d_bytes=int(h2d[0].item()), d2h_bytes=int(h2d[1].item()))
    lib.lib.treeqp_b200_set_profiling(2)
    sh.solve_resident()
    prof = sh.stats()
    lib.lib.treeqp_b200_set_profiling(1)
    sh.solve_resident()
    prof["serialised"] = sh.stats()["kernels"]
    lib.lib.treeqp_b200_set_profiling(0)
    comm_ms = prof["kernels"].get("nccl+pack", (0.0, 0))
    shard = {"cut_depth": plan["cut_depth"], "cut_nodes": plan["ncut"], "local_nodes": plan["n_local"],
             "collectives_ms_per_solve": comm_ms[0], "collectives_per_solve": comm_ms[1],
             "us_per_collective": 1e3 * comm_ms[0] / max(comm_ms[1], 1)}
    sh.destroy()
    comm.destroy()
    return dict(t_res=t_res, t_e2e=t_e2e, dev_ms=dev_ms, stats=stats, e2e_stats=e2e_stats, prof=prof, sol=sol,
                kkt=None, kkt_dev=None, clk=(clk_lo, clk_hi), launches=stats["kernel_launches"] * args.steps * world, shard=shard,
                pinned=pinned)


def roofline_block(prob, prof, world, depth):
    hbm, hbm_src, fp64, fp64_src = measured_peaks()
    alg = algorithmic_bytes(prob)
    kern = prof["kernels"]
    fused = "k_build_factor" in kern
    cb = class_bytes(alg, fused)
    if fused and "k_lambda_step" in kern and "k_stage_fused<0>" in kern and kern["k_stage_fused<0>"][1] < kern["k_lambda_step"][1]:
        # large trees: k_step_affine (lambda step + next iteration's stage solution from the affine model, vectors only)
        # on the iterations that are not refreshed by k_stage_fused<0>, a plain lambda step on the others
        n_step, n_refresh = kern["k_lambda_step"][1], max(kern["k_stage_fused<0>"][1] - 2, 0)  # minus iteration 0 and the final clip-only pass
        b_aff = 8 * int((10 * prob.nx + 10 * prob.nu + 2).sum())
        b_plain = 8 * int(3 * prob.nx.sum())
        cb["k_lambda_step"] = ((n_step - n_refresh) * b_aff + n_refresh * b_plain) / max(n_step, 1)
    lpc = launches_per_call(prob, cb)
    top = max((k for k in kern if k in cb), key=lambda k: kern[k][0])
    t_ms, n_launch = kern[top]
    calls = n_launch / lpc[top]
    share = 1.0 / world           # a rank of a sharded solve moves its subtrees' share of the bytes
    achieved = cb[top] * share * calls / (t_ms * 1e-3) / 1e9
    per_kernel = {}
    for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0]):
        rec = {"ms": round(v[0], 4), "launches": v[1]}
        if k in cb:
            rec["GBps"] = round(cb[k] * share * (v[1] / lpc[k]) / (v[0] * 1e-3) / 1e9, 1)
            rec["frac"] = round(rec["GBps"] / hbm, 3)
        per_kernel[k] = rec
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
        if tr.get("depth") == depth and world == 1 and top in tr["per_launch_bytes"]:
            traffic, traffic_src = tr["per_launch_bytes"][top], tr["source"]
    except Exception:
        pass
    avg_us = 1e3 * t_ms / n_launch
    out = {"bound": "hbm", "kernel": top + (" (fused K2+K3 of SURVEY 8d, one launch per depth)" if top == "k_build_factor" else ""),
           "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
           "traffic": traffic, "traffic_source": traffic_src,
           "frac_dram": (traffic / (avg_us * 1e-6) / 1e9 / hbm) if traffic else None,
           "algorithmic_bytes_per_launch": cb[top] * share / lpc[top], "peak_source": hbm_src,
           "algorithmic_bytes_per_call": cb[top] * share, "launches_per_call": lpc[top],
           "avg_launch_us": avg_us, "kernel_share_of_profiled_solve": t_ms / prof["gpu_ms"],
           "note": "CUDA events on the solver's stream inside one extra solve: a pair around every launch, except that the "
                   "launches of one level sweep (k_build_factor / k_backsolve_fused, one launch per depth, chained with "
                   "programmatic dependent launch) share ONE pair -- avg_launch_us = sweep time / launches of the sweep, "
                   "measured with the overlap the solver runs with; per_kernel_serialised = a pair around every single launch.  "
                   "frac credits k_build_factor with the SURVEY 8d contract bytes (K2 + K3: Hessian blocks written and re-read), which the fused kernel never "
                   "moves and which the re-use of unchanged blocks (27 of 29 iterations redo only right-hand sides on the wide levels) does not even "
                   "compute: frac_dram (measured DRAM bytes of a whole solve over the same time) is the fraction of the memory system actually used",
           "per_kernel": per_kernel,
           "per_kernel_serialised": {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in sorted(prof.get("serialised", {}).items(), key=lambda kv: -kv[1][0])}}
    if top in ("k_build_factor", "k_factor_level"):
        tf = alg["factor_flops"] * share * calls / (t_ms * 1e-3) / 1e12
        out["compute"] = {"achieved": tf, "peak": fp64, "unit": "TFLOP/s", "frac": tf / fp64, "peak_source": fp64_src,
                          "flops_per_call": alg["factor_flops"] * share}
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
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (config 4 batch, config 2 SDUNES, ...)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    # host threads of the packer (OpenMP): torchrun exports OMP_NUM_THREADS=1 to every rank, which
    # would serialise the BLASFEO -> SoA conversion of the e2e path; give each rank its share of cores
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    if world > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, cores // world))
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
    prob = workload(args.depth, 0)

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

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    if world == 1:
        R = run_single(lib, args, prob, barrier, max_over_ranks, clocks, rank)
    else:
        R = run_sharded(lib, args, prob, rank, world, local, dist, barrier, max_over_ranks, clocks)
    clk = clocks.stop(*R["clk"]) if rank == 0 else None

    extra = {}
    if dist is not None and not args.no_extra:
        # replicas: every rank solves its own copy of the tree, no collective (weak scaling of independent QPs)
        try:
            rep = run_single(lib, args, prob, barrier, max_over_ranks, ClockSampler(local), rank)
            extra["replicas_one_tree_per_gpu"] = {
                "workload": "every rank solves its own copy of the N=1 tree, no data-path collective", "scaling": "weak",
                "solves_per_s_resident": world * args.steps / rep["t_res"], "solves_per_s_e2e": world * args.steps / rep["t_e2e"],
                "ms_per_solve_resident": 1e3 * rep["t_res"] / args.steps, "ms_per_solve_e2e": 1e3 * rep["t_e2e"] / args.steps,
                "pack_ms": rep["e2e_stats"]["pack_ms"], "h2d_ms": rep["e2e_stats"]["h2d_ms"]}
        except Exception as e:  # report, never hide
            extra["replicas_one_tree_per_gpu"] = {"error": repr(e)}
        try:
            extra["batch4096_linear_chain"] = bench_batch(lib, 4096, 3, rank, world, barrier, max_over_ranks)
        except Exception as e:
            extra["batch4096_linear_chain"] = {"error": repr(e)}

    if rank != 0:
        if dist is not None:
            dist.barrier()      # rank 0 runs the reference for the parity block meanwhile
            dist.destroy_process_group()
        return

    iters = R["stats"]["iter"]
    roofline = roofline_block(prob, R["prof"], world, args.depth)

    # ---- parity against the unmodified reference on the SAME QP (+ the CPU baseline at N = 1)
    cpu_baseline, parity = None, None
    if ref_lib_path(True) and not args.no_cpu_baseline:
        threads = omp_threads(cores)
        once = run_reference_solve(prob)
        nrep = 5 if world == 1 else 1
        runs = [once() for _ in range(nrep)]
        t_cpu = min(r[0] for r in runs)
        _, st_cpu, it_cpu = runs[-1]
        parity = parity_block(prob, R["sol"], out_dict(once.qp_out, st_cpu),
                              "unmodified reference (oracle/_ref, treeqp_tdunes_solve) on the same QP, %s OpenMP threads" % threads)
        if world > 1:
            parity["owners_per_node_ok"] = R["sol"]["owners_per_node_ok"]
            parity["ok"] = bool(parity["ok"] and R["sol"]["owners_per_node_ok"])
        parity["ls_trials"] = R["stats"]["ls_total"]
        if world == 1:
            cpu_baseline = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": threads or cores, "host_cores": cores, "kind": "reference",
                            "sample": "min over %d full solves of the same QP (%d Newton iterations, status %d), OpenMP build of the unmodified reference" % (nrep, it_cpu, st_cpu),
                            "us_per_newton_iter": 1e6 * t_cpu / max(it_cpu, 1), "newton_iters": it_cpu,
                            "ms_per_solve_all": [round(1e3 * r[0], 1) for r in runs]}
        omp_threads(max(1, cores // world))
    elif not args.no_cpu_baseline:
        parity = {"ok": None, "against": "oracle/_ref missing: parity unchecked in this run"}
        if world == 1:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    if dist is not None:
        dist.barrier()

    if world == 1 and not args.no_extra:
        for name, fn in (("batch4096_linear_chain", bench_batch), ("sdunes_spring_mass_cfg2", bench_sdunes),
                         ("dense_stage_qps_cfg5", bench_dense), ("mpc_resolve_x0_fast_path", bench_mpc)):
            try:
                extra[name] = fn(lib)
            except Exception as e:
                extra[name] = {"error": repr(e)}

    t_res, t_e2e, es = R["t_res"], R["t_e2e"], R["e2e_stats"]
    out = {
        "metric": METRIC, "value": args.steps / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args.depth, prob.Nn, world),
        "detail": {"newton_iters": iters, "ls_trials": R["stats"]["ls_total"], "kkt": R["kkt"], "kkt_device": R["kkt_dev"],
                   "parallelism": "1 QP on 1 GPU" if world == 1 else "1 QP sharded by subtree over %d GPUs: NCCL all-gather of the cut-node Schur complements + all-reduced scalars" % world,
                   "sharding": R.get("shard")},
        "us_per_newton_iter": 1e3 * (1e3 * t_res / args.steps) / max(iters, 1),
        "device_ms_per_step": R["dev_ms"] / args.steps,
        "e2e": {"value": args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": es["h2d_bytes"],
                "d2h_bytes_per_step": es["d2h_bytes"], "ms_per_step": 1e3 * t_e2e / args.steps,
                "pack_ms": es["pack_ms"], "h2d_ms": es["h2d_ms"], "d2h_ms": es["d2h_ms"], "gpu_ms": es["gpu_ms"],
                "host_buffers": ("caller's tree_qp_in / tree_qp_out page-locked once with treeqp_b200_host_register: DMA straight from / into them"
                                 if R.get("pinned") else "pageable: every byte through the host packer and the library's pinned staging block"),
                "ms_per_step_unpinned_buffers": (1e3 * R["t_e2e_unpinned"] if R.get("t_e2e_unpinned") else None)},
        "gpu_launches": int(R["launches"]), "clocks": clk, "parity": parity, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "other_workloads": extra,
    }
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and parity.get("ok") is False:
        sys.stderr.write("bench.py: PARITY FAILURE against the reference: %s\n" % json.dumps(parity))
        raise SystemExit(1)


if __name__ == "__main__":
    main()

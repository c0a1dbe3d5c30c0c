"""The drop-in boundary exercised the way a treeQP user would: the reference's OWN example drivers
(examples/spring_mass_dual_newton_tree.c, ..._scenarios.c), UNMODIFIED, compiled against the reference
headers and linked with the line INTEGRATION.md documents (oracle/Makefile `dropin`: -ltreeqp_b200 first,
then the reference's static archives).

CPU: every dual Newton symbol of the executable must be undefined (= bound to libtreeqp_b200.so at load
time) and no dual_Newton_*.o may have been pulled out of libtreeqp.a -- otherwise the program would run
the reference's CPU solver without anyone noticing.  GPU: the binaries run (NREP = 20 solves each, the
examples' own KKT < 1e-8 asserts are live), the library reports kernel launches, and the files written by
write_solution_to_txt / write_scenarios_solution_to_txt hold the golden solution."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases as CS
from treeqp_b200 import problems as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
TREE = os.path.join(DROPIN, "spring_mass_dual_newton_tree")
SCEN = os.path.join(DROPIN, "spring_mass_dual_newton_scenarios")

SOLVER_SYMS = {
    TREE: ["treeqp_tdunes_opts_calculate_size", "treeqp_tdunes_opts_create", "treeqp_tdunes_opts_set_default",
           "treeqp_tdunes_calculate_size", "treeqp_tdunes_create", "treeqp_tdunes_set_dual_initialization",
           "treeqp_tdunes_solve", "write_solution_to_txt"],
    SCEN: ["treeqp_sdunes_calculate_size", "treeqp_sdunes_create", "treeqp_sdunes_opts_set_default",
           "treeqp_sdunes_set_dual_initialization", "treeqp_sdunes_solve", "treeqp_sdunes_calculate_dual_dimension",
           "write_scenarios_solution_to_txt"],
}


def _build():
    if os.path.isdir("/root/reference/examples"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref", "dropin", "-j8"])
    if not (os.path.exists(TREE) and os.path.exists(SCEN)):
        pytest.skip("oracle/_ref/dropin not built (needs /root/reference at build time)")


@pytest.mark.parametrize("exe", [TREE, SCEN])
def test_examples_bind_to_the_gpu_library(exe):
    _build()
    nm = subprocess.check_output(["nm", exe], text=True)
    table = {ln.split()[-1]: ln.split()[-2] for ln in nm.splitlines() if len(ln.split()) >= 2}
    for sym in SOLVER_SYMS[exe]:
        assert table.get(sym) == "U", "%s is %r in %s: the CPU solver was linked in" % (sym, table.get(sym), os.path.basename(exe))
    # no solver object of the reference inside the executable
    link_map = open(exe + ".map").read()
    assert not re.search(r"libtreeqp\.a\(dual_Newton", link_map), "a dual_Newton_*.o was pulled out of libtreeqp.a"
    needed = subprocess.check_output(["readelf", "-d", exe], text=True)
    assert "libtreeqp_b200.so" in needed
    # and the product library really exports them
    prod = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "treeqp_b200", "lib", "libtreeqp_b200.so")], text=True)
    exported = {ln.split()[-1] for ln in prod.splitlines() if ln.split()}
    assert set(SOLVER_SYMS[exe]) <= exported


def _workdir(tmp_path):
    """cwd of the examples: they read their start vectors from examples/spring_mass_utils/*.txt
    (zeros and x0 in the reference tree, spring_mass_utils/{lambda0_tree,mu0_scen,lambda0_scen,x0}.txt)"""
    d = tmp_path / "examples" / "spring_mass_utils"
    d.mkdir(parents=True)
    data = np.load(os.path.join(ROOT, "tests", "golden", "spring_mass_data.npz"))
    np.savetxt(d / "x0.txt", data["x0"], fmt="%.16e")
    for name in ("lambda0_tree.txt", "mu0_scen.txt", "lambda0_scen.txt"):
        np.savetxt(d / name, np.zeros(4096), fmt="%.16e")
    return d


def _run(exe, tmp_path):
    env = dict(os.environ, TREEQP_B200_VERBOSE="1")
    r = subprocess.run([exe], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stderr.splitlines() if ln.startswith("[treeqp_b200] treeqp_")]
    assert len(lines) == 20, "the example makes NREP = 20 solves; the GPU library saw %d" % len(lines)
    m = re.search(r"status (\d+), (\d+) Newton iterations, (\d+) kernel launches", lines[-1])
    assert m and int(m.group(1)) == 0 and int(m.group(3)) > 0
    return int(m.group(2)), r.stdout


@pytest.mark.gpu
def test_tree_example_runs_on_the_gpu(tmp_path):
    if not os.path.exists(TREE):
        pytest.skip("oracle/_ref/dropin not built")
    d = _workdir(tmp_path)
    iters, out = _run(TREE, tmp_path)
    assert "Maximum error in KKT residuals (tdunes)" in out
    g = CS.golden("spring_mass_cfg1")
    assert iters == int(g["iter"]) == int(np.loadtxt(d / "iter.txt"))
    assert CS.rel_err(np.loadtxt(d / "x_opt.txt"), g["x"]) < 1e-10
    assert CS.rel_err(np.loadtxt(d / "u_opt.txt"), g["u"]) < 1e-10
    assert CS.rel_err(np.loadtxt(d / "lambda_opt.txt"), g["lam"]) < 1e-10


@pytest.mark.gpu
def test_scenarios_example_runs_on_the_gpu(tmp_path):
    if not os.path.exists(SCEN):
        pytest.skip("oracle/_ref/dropin not built")
    d = _workdir(tmp_path)
    iters, out = _run(SCEN, tmp_path)
    assert int(np.loadtxt(d / "iter.txt")) == iters
    data = np.load(os.path.join(ROOT, "tests", "golden", "spring_mass_data.npz"))
    md, Nr, Nh, NX, NU = (int(data[k]) for k in ("md", "Nr", "Nh", "NX", "NU"))
    Ns = md ** Nr
    x = np.loadtxt(d / "xIter.txt").reshape(Ns, Nh, NX)
    u = np.loadtxt(d / "uIter.txt").reshape(Ns, Nh, NU)
    # scenario iterates agree where scenarios share a node (non-anticipativity) and respect the bounds
    assert np.abs(u[:, 0, :] - u[0, 0, :]).max() < 1e-8
    assert np.all(u <= 1.0 + 1e-12) and np.all(u >= -1.0 - 1e-12)
    assert np.isfinite(x).all() and np.isfinite(np.loadtxt(d / "muIter.txt")).all()


# ---- PROFILE = 3 callers (Makefile.rule:112-113): the workspace structs carry the reference's `timings` member and the drivers
# print it.  They link libtreeqp_b200_profile3.so (same solvers, PROFILE = 3 struct layout, timers fed from CUDA events).
TREE_P3, SCEN_P3 = TREE + "_p3", SCEN + "_p3"


@pytest.mark.parametrize("exe", [TREE_P3, SCEN_P3])
def test_profile3_examples_bind_to_the_profile3_library(exe):
    _build()
    if not os.path.exists(exe):
        pytest.skip("PROFILE = 3 drop-in not built")
    nm = subprocess.check_output(["nm", exe], text=True)
    table = {ln.split()[-1]: ln.split()[-2] for ln in nm.splitlines() if len(ln.split()) >= 2}
    base = TREE if exe == TREE_P3 else SCEN
    for sym in SOLVER_SYMS[base] + ["timers_print"]:
        assert table.get(sym) == "U", (sym, table.get(sym))
    assert "libtreeqp_b200_profile3.so" in subprocess.check_output(["readelf", "-d", exe], text=True)
    assert not re.search(r"libtreeqp\.a\((dual_Newton|profiling)", open(exe + ".map").read())
    # the default library, linked by a PROFILE > 0 caller by mistake, stops the program instead of printing garbage
    prod = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "treeqp_b200", "lib", "libtreeqp_b200_profile3.so")], text=True)
    exported = {ln.split()[-1] for ln in prod.splitlines() if ln.split()}
    assert {"timers_print", "timers_write_to_txt", "timers_create", "timers_update", "timers_calculate_size"} <= exported


def test_profile3_struct_layout_matches_reference():
    """sizeof of the PROFILE = 3 workspaces and the offset of `timings` as gcc sees them through our header and through the
    reference's: the caller's compiled-in layout is what the library writes into"""
    if not os.path.isdir("/root/reference/treeqp"):
        pytest.skip("needs the reference headers")
    import tempfile
    src_ours = '#include <stdio.h>\n#include <stddef.h>\n#include "treeqp_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(treeqp_tdunes_workspace), offsetof(treeqp_tdunes_workspace, timings), sizeof(treeqp_sdunes_workspace), offsetof(treeqp_sdunes_workspace, timings), sizeof(treeqp_profiling_t));return 0;}\n'
    src_ref = '#include <stdio.h>\n#include <stddef.h>\n#include "treeqp/src/dual_Newton_tree.h"\n#include "treeqp/src/dual_Newton_scenarios.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(treeqp_tdunes_workspace), offsetof(treeqp_tdunes_workspace, timings), sizeof(treeqp_sdunes_workspace), offsetof(treeqp_sdunes_workspace, timings), sizeof(treeqp_profiling_t));return 0;}\n'
    outs = []
    with tempfile.TemporaryDirectory() as td:
        for name, src, inc in (("ours", src_ours, ["-I" + os.path.join(ROOT, "include")]),
                               ("ref", src_ref, ["-I/root/reference", "-I/root/reference/external", "-I/root/reference/external/blasfeo/include",
                                                 "-I" + os.path.join(ROOT, "oracle", "_ref", "include")])):
            c = os.path.join(td, name + ".c")
            open(c, "w").write(src)
            exe = os.path.join(td, name)
            subprocess.check_call(["gcc", "-DPROFILE=3", "-w", c, "-o", exe] + inc)
            outs.append(subprocess.check_output([exe], text=True).split())
    assert outs[0] == outs[1], outs


def _run_p3(exe, tmp_path):
    r = subprocess.run([exe], cwd=tmp_path, env=dict(os.environ), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.gpu
def test_profile3_tree_example_prints_gpu_timers(tmp_path):
    if not os.path.exists(TREE_P3):
        pytest.skip("PROFILE = 3 drop-in not built")
    d = _workdir(tmp_path)
    out = _run_p3(TREE_P3, tmp_path)
    assert "Maximum error in KKT residuals (tdunes)" in out          # the example's own KKT assert passed (rc 0)
    g = CS.golden("spring_mass_cfg1")
    m = re.search(r"algorithm converged in \((\d+) it\):\s+([0-9.]+) ms", out)
    assert m and int(m.group(1)) == int(g["iter"]) and float(m.group(2)) > 0.0
    assert out.count("Iteration #") == int(g["iter"])
    for phase in ("solved stage QPs in", "built dual problem in", "calculated Newton direction in", "performed line-search"):
        mm = re.search(re.escape(phase) + r".*?:\s+([0-9.]+) ms", out)
        assert mm and float(mm.group(1)) > 0.0, phase
    # the files the reference's write_solution_to_txt adds under PROFILE = 3 (utils/print.c:312-345)
    ls = np.loadtxt(d / "ls_iters.txt")
    assert ls[:int(g["iter"])].sum() > 0
    st = np.loadtxt(d / "newton_direction_times.txt")
    assert np.all(st[:int(g["iter"])] > 0.0) and np.isnan(st[int(g["iter"])])
    assert CS.rel_err(np.loadtxt(d / "x_opt.txt"), g["x"]) < 1e-10


@pytest.mark.gpu
def test_profile3_scenarios_example_prints_gpu_timers(tmp_path):
    if not os.path.exists(SCEN_P3):
        pytest.skip("PROFILE = 3 drop-in not built")
    _workdir(tmp_path)
    out = _run_p3(SCEN_P3, tmp_path)
    m = re.search(r"algorithm converged in \((\d+) it\):\s+([0-9.]+) ms", out)
    assert m and int(m.group(1)) > 0 and float(m.group(2)) > 0.0
    assert "calculated Newton direction in" in out

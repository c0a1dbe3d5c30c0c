"""CPU tests of the drop-in boundary: the product library loads without a GPU, exports every
symbol include/treeqp_b200.h declares, and its host containers are layout-compatible with the
reference's (a tree_qp_in built by one library is readable by the other)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases as CS
from treeqp_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "treeqp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = set(re.findall(r"\b(treeqp_\w+|tree_qp_\w+|setup_multistage_tree|calculate_number_of_nodes|"
                           r"number_of_nodes_from_nkids)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol(product_lib):
    missing = [n for n in declared_symbols() if not hasattr(product_lib.lib, n)]
    assert not missing, missing
    assert len(declared_symbols()) >= 35


def test_default_options_match_reference(product_lib, ref_lib):
    prob, _ = CS.get_case("spring_mass_cfg1")
    vals = []
    for lib in (product_lib, ref_lib):
        o = capi.TdunesOpts()
        mem = capi._Buf(lib.lib.treeqp_tdunes_opts_calculate_size(prob.Nn))
        lib.lib.treeqp_tdunes_opts_create(prob.Nn, C.byref(o), mem.ptr)
        lib.lib.treeqp_tdunes_opts_set_default(prob.Nn, C.byref(o))
        vals.append([getattr(o, f) for f, _ in capi.TdunesOpts._fields_ if f != "qp_solver"] +
                    [o.qp_solver[k] for k in range(prob.Nn)])
    assert vals[0] == vals[1]


@pytest.mark.parametrize("name", ["spring_mass_cfg1", "random_vary_dims", "linear_chain_cfg4"])
def test_container_layout_is_interchangeable(name, product_lib, ref_lib):
    """Same sizes, and the reference's solver accepts a tree_qp_in carved by OUR containers."""
    prob, opts = CS.get_case(name)
    nk, nx, nu = (np.ascontiguousarray(a, dtype=np.int32) for a in (prob.nk, prob.nx, prob.nu))
    ip = lambda a: a.ctypes.data_as(capi.c_int_p)
    for fn in ("tree_qp_in_calculate_size",):
        assert getattr(product_lib.lib, fn)(prob.Nn, ip(nx), ip(nu), None, ip(nk)) == \
            getattr(ref_lib.lib, fn)(prob.Nn, ip(nx), ip(nu), None, ip(nk))
    assert product_lib.lib.tree_qp_out_calculate_size(prob.Nn, ip(nx), ip(nu), None) == \
        ref_lib.lib.tree_qp_out_calculate_size(prob.Nn, ip(nx), ip(nu), None)

    qp_in = capi.QpIn(product_lib, prob)        # built by our containers.c
    qp_out = capi.QpOut(product_lib, prob)
    qp_in.lib = ref_lib                          # ... solved by the reference's TDUNES
    solver = capi.Tdunes(ref_lib, qp_in, **opts)
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    status = solver.solve(qp_out)
    g = CS.golden(name)
    assert status == int(g["status"]) and qp_out.iter == int(g["iter"])
    assert CS.rel_err(qp_out.x, g["x"]) < 1e-12 and CS.rel_err(qp_out.lam, g["lam"]) < 1e-12
    # both KKT checkers agree on the result
    k_ref = ref_lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr)
    k_own = product_lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr)
    assert k_ref < 1e-8 and k_own < 1e-8 and abs(k_ref - k_own) < 1e-9


def test_tree_helpers_match_reference(product_lib, ref_lib):
    for md, Nr, Nh in [(3, 2, 10), (2, 16, 16), (1, 0, 7), (4, 2, 20), (3, 4, 20)]:
        a, b = product_lib.multistage_tree(md, Nr, Nh), ref_lib.multistage_tree(md, Nr, Nh)
        assert np.array_equal(a, b)


def test_no_gpu_fails_loudly(product_lib):
    """Without a CUDA device the solve must report an error, never fall back to a CPU path."""
    if product_lib.lib.treeqp_b200_device_count() > 0:
        pytest.skip("a GPU is visible")
    prob, _ = CS.get_case("spring_mass_cfg1")
    qp_in, qp_out = capi.QpIn(product_lib, prob), capi.QpOut(product_lib, prob)
    solver = capi.Tdunes(product_lib, qp_in)
    solver.set_dual_initialization(np.zeros(prob.n_lambda))
    assert solver.solve(qp_out) == capi.TREEQP_UNKNOWN_ERROR


def test_recreate_on_same_memory_does_not_leak_contexts(product_lib):
    """The reference API has create but no destroy: a caller that re-creates its workspace over the same memory
    (the benchmark harnesses do, once per QP) must not accumulate device arenas (registry.c).  Without a GPU the
    context still exists (it then holds no arena), so the bookkeeping is testable here."""
    import ctypes as C
    from treeqp_b200 import problems as P
    L = product_lib.lib
    L.treeqp_b200_live_contexts.restype = C.c_int
    prob = P.spring_mass(3, 2, 10)
    qp_in = capi.QpIn(product_lib, prob)
    s = capi.Tdunes(product_lib, qp_in)
    n0 = L.treeqp_b200_live_contexts()
    for _ in range(5):   # the reference's call sequence, again and again on the same work / memory
        L.treeqp_tdunes_create(qp_in.ptr, C.byref(s.opts), s.work.ptr, s.work_mem.ptr)
    assert L.treeqp_b200_live_contexts() == n0
    s.destroy()
    assert L.treeqp_b200_live_contexts() == n0 - 1


def test_host_register_without_gpu_fails_cleanly(product_lib):
    """treeqp_b200_host_register page-locks a caller buffer through the CUDA runtime: without a device it reports
    failure (non-zero) and the caller simply stays on the staged path; unregistering what was never registered is
    harmless."""
    import ctypes as C
    L = product_lib.lib
    if L.treeqp_b200_device_count() > 0:
        pytest.skip("needs a box without a GPU")
    buf = capi._Buf(1 << 21)
    assert L.treeqp_b200_host_register(buf.ptr, C.c_size_t(buf.arr.size)) != 0
    assert L.treeqp_b200_host_unregister(buf.ptr) != 0
    assert L.treeqp_b200_host_unregister(None) == 0
    assert buf.pin(product_lib) is False

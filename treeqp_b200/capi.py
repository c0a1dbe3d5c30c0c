"""ctypes mirror of the treeQP C API (reference: treeqp/src/tree_qp_common.h:139-323,
treeqp/src/dual_Newton_tree.h:154-168, treeqp/utils/tree.h:56-73).

The classes below drive *any* shared library that exports the reference's C symbols:
``libtreeqp_b200.so`` (the product: CUDA kernels behind the C ABI) and, in the tests only, the
unmodified reference compiled to ``oracle/_ref/libtreeqp_ref.so``.  The same Python code builds
the QP through the same setters and calls the same ``treeqp_tdunes_*`` entry points on both, so
a parity test reads like the reference's own examples (examples/spring_mass_dual_newton_tree.c).

Python is harness only: nothing here computes; the arithmetic is in the shared library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.environ.get("TREEQP_B200_LIB", os.path.join(_HERE, "lib", "libtreeqp_b200.so"))

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)

# return codes, treeqp/utils/types.h:55-78
TREEQP_OPTIMAL_SOLUTION_FOUND = 0
TREEQP_MAXIMUM_ITERATIONS_REACHED = 1
TREEQP_DN_NOT_DESCENT_DIRECTION = 2
TREEQP_OK = 7
TREEQP_INVALID_OPTION = 9
TREEQP_UNKNOWN_ERROR = 11
TREEQP_SUMSQUAREDERRORS, TREEQP_TWONORM, TREEQP_INFNORM = 0, 1, 2
TREEQP_CLIPPING_SOLVER, TREEQP_QPOASES_SOLVER = 0, 1
TREEQP_NO_REGULARIZATION, TREEQP_ALWAYS_LM, TREEQP_ON_THE_FLY_LM = 0, 1, 2
TREEQP_INF = 1e12


class TdunesOpts(C.Structure):
    """treeqp_tdunes_opts_t, dual_Newton_tree.h:67-87"""
    _fields_ = [
        ("maxIter", C.c_int),
        ("qp_solver", c_int_p),
        ("checkLastActiveSet", C.c_int),
        ("stationarityTolerance", C.c_double),
        ("termCondition", C.c_int),
        ("regType", C.c_int),
        ("regTol", C.c_double),
        ("regValue", C.c_double),
        ("lineSearchMaxIter", C.c_int),
        ("lineSearchGamma", C.c_double),
        ("lineSearchBeta", C.c_double),
        ("lineSearchRestartTrigger", C.c_int),
    ]


class SdunesOpts(C.Structure):
    """treeqp_sdunes_opts_t, dual_Newton_scenarios.h:49-66"""
    _fields_ = [
        ("maxIter", C.c_int),
        ("checkLastActiveSet", C.c_int),
        ("stationarityTolerance", C.c_double),
        ("termCondition", C.c_int),
        ("regType", C.c_int),
        ("regTol", C.c_double),
        ("regValue", C.c_double),
        ("lineSearchMaxIter", C.c_int),
        ("lineSearchGamma", C.c_double),
        ("lineSearchBeta", C.c_double),
    ]


class DVec(C.Structure):
    """struct blasfeo_dvec (HIGH_PERFORMANCE), blasfeo_common.h:72-78"""
    _fields_ = [("m", C.c_int), ("pm", C.c_int), ("pa", c_double_p), ("memsize", C.c_int)]


class Info(C.Structure):
    """treeqp_info_t, tree_qp_common.h:43-51"""
    _fields_ = [("Nn", C.c_int), ("iter", C.c_int), ("total_time", C.c_double),
                ("solver_time", C.c_double), ("interface_time", C.c_double)]


class QpOutStruct(C.Structure):
    """tree_qp_out, tree_qp_common.h:121-133"""
    _fields_ = [("info", Info), ("x", C.POINTER(DVec)), ("u", C.POINTER(DVec)),
                ("lam", C.POINTER(DVec)), ("mu_x", C.POINTER(DVec)), ("mu_u", C.POINTER(DVec)),
                ("mu_d", C.POINTER(DVec))]


class Stats(C.Structure):
    """treeqp_b200_stats (include/treeqp_b200.h, additive)"""
    _fields_ = [("iter", C.c_int), ("ls_total", C.c_int), ("ls_iters", C.c_int * 256),
                ("n_reg", C.c_int), ("gpu_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("pack_ms", C.c_double), ("h2d_bytes", C.c_longlong),
                ("d2h_bytes", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("kernel_ms", C.c_double * 16), ("kernel_count", C.c_int * 16)]


_DVEC_DT = np.dtype([("m", "<i4"), ("pm", "<i4"), ("pa", "<u8"), ("memsize", "<i4"), ("_pad", "<i4")])
assert _DVEC_DT.itemsize == C.sizeof(DVec)


@dataclass
class Problem:
    """A tree QP in flat column-major form (what the reference's ``set_ltv_*`` setters take).

    nk/nx/nu per node (breadth-first); A,B,b concatenated per edge (edge e joins node e+1 to its
    parent); Q,R either diagonals (``Qd``/``Rd``, clipping path) or dense blocks (``Q``/``R``/``S``
    concatenated column-major, dense stage-QP path); q,r,bounds concatenated per node.
    """
    nk: np.ndarray
    nx: np.ndarray
    nu: np.ndarray
    A: np.ndarray
    B: np.ndarray
    b: np.ndarray
    q: np.ndarray
    r: np.ndarray
    xmin: np.ndarray
    xmax: np.ndarray
    umin: np.ndarray
    umax: np.ndarray
    Qd: Optional[np.ndarray] = None
    Rd: Optional[np.ndarray] = None
    Q: Optional[np.ndarray] = None
    R: Optional[np.ndarray] = None
    S: Optional[np.ndarray] = None
    lam0: Optional[np.ndarray] = None
    # general stage constraints dmin <= C x + D u <= dmax (tree_qp_common.h:104-109): nc per node,
    # C (nc x nx) and D (nc x nu) column-major and concatenated per node, dmin/dmax concatenated
    nc: Optional[np.ndarray] = None
    C: Optional[np.ndarray] = None
    D: Optional[np.ndarray] = None
    dmin: Optional[np.ndarray] = None
    dmax: Optional[np.ndarray] = None
    meta: dict = field(default_factory=dict)

    @property
    def Nn(self) -> int:
        return int(len(self.nk))

    @property
    def dad(self) -> np.ndarray:
        dad = np.full(self.Nn, -1, dtype=np.int64)
        first = np.concatenate([[1], 1 + np.cumsum(self.nk)[:-1]])
        for i in np.nonzero(self.nk)[0]:
            dad[first[i]:first[i] + self.nk[i]] = i
        return dad

    @property
    def n_lambda(self) -> int:
        return int(self.nx.sum() - self.nx[0])


def _ip(a):
    return a.ctypes.data_as(c_int_p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class TreeQpLib:
    """A loaded shared library exporting the treeQP C API."""

    def __init__(self, path: str = PRODUCT_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        self.is_product = hasattr(L, "treeqp_b200_version")
        vp = C.c_void_p
        L.tree_qp_in_calculate_size.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
        L.tree_qp_in_calculate_size.restype = C.c_int
        L.tree_qp_in_create.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, vp, vp]
        L.tree_qp_in_create.restype = None
        L.tree_qp_out_calculate_size.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p]
        L.tree_qp_out_calculate_size.restype = C.c_int
        L.tree_qp_out_create.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p, vp, vp]
        L.tree_qp_out_create.restype = None
        L.tree_qp_in_set_ltv_dynamics_colmajor.argtypes = [c_double_p, c_double_p, c_double_p, vp]
        L.tree_qp_in_set_ltv_dynamics_colmajor.restype = None
        L.tree_qp_in_set_ltv_objective_diag.argtypes = [c_double_p] * 4 + [vp]
        L.tree_qp_in_set_ltv_objective_diag.restype = None
        L.tree_qp_in_set_ltv_bounds.argtypes = [c_double_p] * 4 + [vp]
        L.tree_qp_in_set_ltv_bounds.restype = None
        L.tree_qp_in_set_node_objective_colmajor.argtypes = [c_double_p] * 5 + [vp, C.c_int]
        L.tree_qp_in_set_node_objective_colmajor.restype = None
        L.tree_qp_in_set_node_general_constraints.argtypes = [c_double_p] * 4 + [vp, C.c_int]
        L.tree_qp_in_set_node_general_constraints.restype = None
        L.tree_qp_out_max_KKT_res.argtypes = [vp, vp]
        L.tree_qp_out_max_KKT_res.restype = C.c_double
        L.treeqp_tdunes_opts_calculate_size.argtypes = [C.c_int]
        L.treeqp_tdunes_opts_calculate_size.restype = C.c_int
        L.treeqp_tdunes_opts_create.argtypes = [C.c_int, C.POINTER(TdunesOpts), vp]
        L.treeqp_tdunes_opts_create.restype = None
        L.treeqp_tdunes_opts_set_default.argtypes = [C.c_int, C.POINTER(TdunesOpts)]
        L.treeqp_tdunes_opts_set_default.restype = None
        L.treeqp_tdunes_calculate_size.argtypes = [vp, C.POINTER(TdunesOpts)]
        L.treeqp_tdunes_calculate_size.restype = C.c_int
        L.treeqp_tdunes_create.argtypes = [vp, C.POINTER(TdunesOpts), vp, vp]
        L.treeqp_tdunes_create.restype = None
        L.treeqp_tdunes_set_dual_initialization.argtypes = [c_double_p, vp]
        L.treeqp_tdunes_set_dual_initialization.restype = None
        L.treeqp_tdunes_solve.argtypes = [vp, vp, C.POINTER(TdunesOpts), vp]
        L.treeqp_tdunes_solve.restype = C.c_int
        # SDUNES septet + dual dimension, dual_Newton_scenarios.h:143-163
        L.treeqp_sdunes_opts_calculate_size.argtypes = [C.c_int]
        L.treeqp_sdunes_opts_calculate_size.restype = C.c_int
        L.treeqp_sdunes_opts_create.argtypes = [C.c_int, C.POINTER(SdunesOpts), vp]
        L.treeqp_sdunes_opts_create.restype = None
        L.treeqp_sdunes_opts_set_default.argtypes = [C.c_int, C.POINTER(SdunesOpts)]
        L.treeqp_sdunes_opts_set_default.restype = None
        L.treeqp_sdunes_calculate_size.argtypes = [vp, C.POINTER(SdunesOpts)]
        L.treeqp_sdunes_calculate_size.restype = C.c_int
        L.treeqp_sdunes_create.argtypes = [vp, C.POINTER(SdunesOpts), vp, vp]
        L.treeqp_sdunes_create.restype = None
        L.treeqp_sdunes_calculate_dual_dimension.argtypes = [C.c_int, C.c_int, C.c_int]
        L.treeqp_sdunes_calculate_dual_dimension.restype = C.c_int
        L.treeqp_sdunes_set_dual_initialization.argtypes = [c_double_p, c_double_p, vp]
        L.treeqp_sdunes_set_dual_initialization.restype = None
        L.treeqp_sdunes_solve.argtypes = [vp, vp, C.POINTER(SdunesOpts), vp]
        L.treeqp_sdunes_solve.restype = C.c_int
        L.tree_qp_in_eliminate_x0.argtypes = [vp]
        L.tree_qp_in_eliminate_x0.restype = None
        L.tree_qp_out_eliminate_x0.argtypes = [vp]
        L.tree_qp_out_eliminate_x0.restype = None
        L.tree_qp_in_set_x0_colmaj.argtypes = [vp, c_double_p]
        L.tree_qp_in_set_x0_colmaj.restype = None
        L.setup_multistage_tree.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p]
        L.setup_multistage_tree.restype = None
        L.calculate_number_of_nodes.argtypes = [C.c_int, C.c_int, C.c_int]
        L.calculate_number_of_nodes.restype = C.c_int
        if self.is_product:
            L.treeqp_b200_version.restype = C.c_char_p
            L.treeqp_b200_host_register.argtypes = [vp, C.c_size_t]
            L.treeqp_b200_host_register.restype = C.c_int
            L.treeqp_b200_host_unregister.argtypes = [vp]
            L.treeqp_b200_host_unregister.restype = C.c_int
            L.treeqp_b200_live_contexts.restype = C.c_int
            L.treeqp_b200_device_count.restype = C.c_int
            L.treeqp_b200_set_device.argtypes = [C.c_int]
            L.treeqp_b200_set_device.restype = C.c_int
            L.treeqp_tdunes_update_x0.argtypes = [vp, vp]
            L.treeqp_tdunes_update_x0.restype = C.c_int
            L.treeqp_tdunes_warm_start_resident.argtypes = [vp]
            L.treeqp_tdunes_warm_start_resident.restype = C.c_int
            L.treeqp_tdunes_kkt_resident.argtypes = [vp, c_double_p]
            L.treeqp_tdunes_kkt_resident.restype = C.c_int
            L.treeqp_tdunes_batch_kkt_resident.argtypes = [vp, c_double_p]
            L.treeqp_tdunes_batch_kkt_resident.restype = C.c_int
            L.treeqp_b200_set_profiling.argtypes = [C.c_int]
            L.treeqp_b200_set_profiling.restype = None
            L.treeqp_b200_num_kernel_classes.restype = C.c_int
            L.treeqp_b200_kernel_class_name.argtypes = [C.c_int]
            L.treeqp_b200_kernel_class_name.restype = C.c_char_p
            L.treeqp_tdunes_destroy.argtypes = [vp]
            L.treeqp_tdunes_destroy.restype = None
            L.treeqp_tdunes_get_stats.argtypes = [vp, C.POINTER(Stats)]
            L.treeqp_tdunes_get_stats.restype = None
            L.treeqp_tdunes_upload.argtypes = [vp, C.POINTER(TdunesOpts), vp]
            L.treeqp_tdunes_upload.restype = C.c_int
            L.treeqp_tdunes_solve_resident.argtypes = [C.POINTER(TdunesOpts), vp]
            L.treeqp_tdunes_solve_resident.restype = C.c_int
            L.treeqp_tdunes_download.argtypes = [vp, vp]
            L.treeqp_tdunes_download.restype = C.c_int
            L.treeqp_tdunes_batch_create.argtypes = [C.c_int, C.POINTER(vp), C.POINTER(TdunesOpts)]
            L.treeqp_tdunes_batch_create.restype = vp
            L.treeqp_tdunes_batch_set_dual_initialization.argtypes = [vp, C.c_int, c_double_p]
            L.treeqp_tdunes_batch_set_dual_initialization.restype = None
            L.treeqp_tdunes_batch_upload.argtypes = [vp, C.POINTER(vp)]
            L.treeqp_tdunes_batch_upload.restype = C.c_int
            L.treeqp_tdunes_batch_update_x0.argtypes = [vp, C.POINTER(vp)]
            L.treeqp_tdunes_batch_update_x0.restype = C.c_int
            L.treeqp_tdunes_batch_solve_resident.argtypes = [vp, c_int_p]
            L.treeqp_tdunes_batch_solve_resident.restype = C.c_int
            L.treeqp_tdunes_batch_download.argtypes = [vp, C.POINTER(vp)]
            L.treeqp_tdunes_batch_download.restype = C.c_int
            L.treeqp_tdunes_solve_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), c_int_p]
            L.treeqp_tdunes_solve_batch.restype = C.c_int
            L.treeqp_tdunes_batch_get_stats.argtypes = [vp, C.POINTER(Stats)]
            L.treeqp_tdunes_batch_get_stats.restype = None
            L.treeqp_tdunes_batch_destroy.argtypes = [vp]
            L.treeqp_tdunes_batch_destroy.restype = None
            L.treeqp_b200_comm_unique_id.argtypes = [C.c_char_p]
            L.treeqp_b200_comm_unique_id.restype = C.c_int
            L.treeqp_b200_comm_create.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
            L.treeqp_b200_comm_create.restype = vp
            L.treeqp_b200_comm_destroy.argtypes = [vp]
            L.treeqp_b200_comm_destroy.restype = None
            L.treeqp_tdunes_sharded_create.argtypes = [vp, C.POINTER(TdunesOpts), vp]
            L.treeqp_tdunes_sharded_create.restype = vp
            L.treeqp_tdunes_sharded_set_dual_initialization.argtypes = [vp, c_double_p, vp]
            L.treeqp_tdunes_sharded_set_dual_initialization.restype = None
            L.treeqp_tdunes_sharded_upload.argtypes = [vp, vp]
            L.treeqp_tdunes_sharded_upload.restype = C.c_int
            L.treeqp_tdunes_sharded_solve_resident.argtypes = [vp]
            L.treeqp_tdunes_sharded_solve_resident.restype = C.c_int
            L.treeqp_tdunes_sharded_download.argtypes = [vp, vp]
            L.treeqp_tdunes_sharded_download.restype = C.c_int
            L.treeqp_tdunes_sharded_solve.argtypes = [vp, vp, vp]
            L.treeqp_tdunes_sharded_solve.restype = C.c_int
            L.treeqp_tdunes_sharded_get_stats.argtypes = [vp, C.POINTER(Stats)]
            L.treeqp_tdunes_sharded_get_stats.restype = None
            L.treeqp_tdunes_sharded_get_plan.argtypes = [vp, c_int_p, c_int_p, c_int_p, c_int_p]
            L.treeqp_tdunes_sharded_get_plan.restype = None
            L.treeqp_tdunes_sharded_destroy.argtypes = [vp]
            L.treeqp_tdunes_sharded_destroy.restype = None
            L.treeqp_b200_shard_plan.argtypes = [vp, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]
            L.treeqp_b200_shard_plan.restype = C.c_int
            L.treeqp_sdunes_destroy.argtypes = [vp]
            L.treeqp_sdunes_destroy.restype = None
            L.treeqp_sdunes_get_stats.argtypes = [vp, C.POINTER(Stats)]
            L.treeqp_sdunes_get_stats.restype = None
            L.treeqp_sdunes_upload.argtypes = [vp, C.POINTER(SdunesOpts), vp]
            L.treeqp_sdunes_upload.restype = C.c_int
            L.treeqp_sdunes_solve_resident.argtypes = [C.POINTER(SdunesOpts), vp]
            L.treeqp_sdunes_solve_resident.restype = C.c_int
            L.treeqp_sdunes_download.argtypes = [vp, vp]
            L.treeqp_sdunes_download.restype = C.c_int
            L.treeqp_sdunes_get_duals.argtypes = [c_double_p, c_double_p, vp]
            L.treeqp_sdunes_get_duals.restype = None
            L.treeqp_sdunes_num_phases.restype = C.c_int
            L.treeqp_sdunes_phase_name.argtypes = [C.c_int]
            L.treeqp_sdunes_phase_name.restype = C.c_char_p

    # -- small helpers mirroring utils/tree.c
    def multistage_tree(self, md: int, Nr: int, Nh: int) -> np.ndarray:
        Nn = self.lib.calculate_number_of_nodes(md, Nr, Nh)
        nk = np.zeros(Nn, dtype=np.int32)
        self.lib.setup_multistage_tree(md, Nr, Nh, _ip(nk))
        return nk


def stats_dict(lib: "TreeQpLib", st: Stats, sdunes: bool = False) -> dict:
    n = min(st.iter, 256)
    if sdunes:  # phases of the single persistent kernel
        nk, name = lib.lib.treeqp_sdunes_num_phases(), lib.lib.treeqp_sdunes_phase_name
    else:
        nk, name = lib.lib.treeqp_b200_num_kernel_classes(), lib.lib.treeqp_b200_kernel_class_name
    kernels = {name(k).decode(): (st.kernel_ms[k], st.kernel_count[k])
               for k in range(nk) if st.kernel_count[k] > 0}
    return dict(iter=st.iter, ls_total=st.ls_total, ls_iters=[st.ls_iters[i] for i in range(n)],
                n_reg=st.n_reg, gpu_ms=st.gpu_ms, h2d_ms=st.h2d_ms, d2h_ms=st.d2h_ms,
                pack_ms=st.pack_ms, h2d_bytes=st.h2d_bytes, d2h_bytes=st.d2h_bytes,
                kernel_launches=st.kernel_launches, kernels=kernels)


class _Buf:
    """64-byte aligned host buffer (the caller-owned memory of the reference's protocol)."""

    def __init__(self, nbytes: int):
        # large buffers: page-aligned, whole pages (so that treeqp_b200_host_register can lock all of it)
        align = 4096 if nbytes >= (1 << 20) else 64
        size = (nbytes + align - 1) // align * align
        self.raw = np.zeros(size + align, dtype=np.uint8)
        off = (-self.raw.ctypes.data) % align
        self.arr = self.raw[off:off + size]
        self.ptr = C.c_void_p(self.arr.ctypes.data)
        self.addr = self.arr.ctypes.data
        self.nbytes = nbytes
        self._pinned_by = None

    def pin(self, lib) -> bool:
        """page-lock this buffer for DMA (treeqp_b200_host_register); undone when the buffer dies"""
        if self._pinned_by is None and lib.is_product and lib.lib.treeqp_b200_host_register(self.ptr, C.c_size_t(self.arr.size)) == 0:
            self._pinned_by = lib
        return self._pinned_by is not None

    def unpin(self):
        if self._pinned_by is not None:
            self._pinned_by.lib.treeqp_b200_host_unregister(self.ptr)
            self._pinned_by = None

    def __del__(self):
        try:
            self.unpin()
        except Exception:
            pass


class QpIn:
    """tree_qp_in built through the library's own setters (tree_qp_common.c:148, 1959-2100)."""

    def __init__(self, lib: TreeQpLib, prob: Problem):
        self.lib, self.prob = lib, prob
        L = lib.lib
        self.nk, self.nx, self.nu = _i32(prob.nk), _i32(prob.nx), _i32(prob.nu)
        self.nc = _i32(prob.nc) if prob.nc is not None else None
        ncp = _ip(self.nc) if self.nc is not None else None
        Nn = prob.Nn
        size = L.tree_qp_in_calculate_size(Nn, _ip(self.nx), _ip(self.nu), ncp, _ip(self.nk))
        assert size > 0, "tree_qp_in size overflows the int return type of the reference API"
        self.mem = _Buf(size)
        self.struct = _Buf(1024)  # opaque tree_qp_in (≈400 B)
        L.tree_qp_in_create(Nn, _ip(self.nx), _ip(self.nu), ncp, _ip(self.nk), self.struct.ptr, self.mem.ptr)
        self.ptr = self.struct.ptr
        self.set_data(prob)

    def pin(self) -> bool:
        """page-lock the tree_qp_in memory: large trees then upload by DMA straight from it"""
        return self.mem.pin(self.lib)

    def set_data(self, prob: Problem):
        L = self.lib.lib
        A, B, b = _f64(prob.A), _f64(prob.B), _f64(prob.b)
        L.tree_qp_in_set_ltv_dynamics_colmajor(_dp(A), _dp(B), _dp(b), self.ptr)
        q, r = _f64(prob.q), _f64(prob.r)
        if prob.Qd is not None:
            Qd, Rd = _f64(prob.Qd), _f64(prob.Rd)
            L.tree_qp_in_set_ltv_objective_diag(_dp(Qd), _dp(Rd), _dp(q), _dp(r), self.ptr)
        else:
            Q, R, S = _f64(prob.Q), _f64(prob.R), _f64(prob.S)
            iq = ir = is_ = ix = iu = 0
            for k in range(prob.Nn):
                nx, nu = int(prob.nx[k]), int(prob.nu[k])
                L.tree_qp_in_set_node_objective_colmajor(
                    _dp(Q[iq:]), _dp(R[ir:]) if nu else None, _dp(S[is_:]) if nu * nx else None,
                    _dp(q[ix:]), _dp(r[iu:]) if nu else None, self.ptr, k)
                iq += nx * nx; ir += nu * nu; is_ += nu * nx; ix += nx; iu += nu
        xmin, xmax, umin, umax = _f64(prob.xmin), _f64(prob.xmax), _f64(prob.umin), _f64(prob.umax)
        L.tree_qp_in_set_ltv_bounds(_dp(xmin), _dp(xmax), _dp(umin), _dp(umax), self.ptr)
        if prob.nc is not None and int(np.sum(prob.nc)) > 0:
            Cm, Dm, dlo, dhi = _f64(prob.C), _f64(prob.D), _f64(prob.dmin), _f64(prob.dmax)
            ic = idd = iv = 0
            for k in range(prob.Nn):
                nx, nu, nc = int(prob.nx[k]), int(prob.nu[k]), int(prob.nc[k])
                if nc > 0:
                    L.tree_qp_in_set_node_general_constraints(
                        _dp(Cm[ic:]) if nx else None, _dp(Dm[idd:]) if nu else None, _dp(dlo[iv:]), _dp(dhi[iv:]), self.ptr, k)
                ic += nc * nx; idd += nc * nu; iv += nc


    def eliminate_x0(self):
        """tree_qp_in_eliminate_x0 (tree_qp_common.c:404): the pinned root state leaves the QP"""
        self.lib.lib.tree_qp_in_eliminate_x0(self.ptr)

    def set_x0(self, x0):
        """tree_qp_in_set_x0_colmaj (tree_qp_common.c:2229): works on pinned and eliminated roots"""
        x0 = _f64(x0)
        self.lib.lib.tree_qp_in_set_x0_colmaj(self.ptr, _dp(x0))


class QpOut:
    """tree_qp_out plus vectorised readers of its BLASFEO vector views."""

    def __init__(self, lib: TreeQpLib, prob: Problem):
        self.lib, self.prob = lib, prob
        L = lib.lib
        nx, nu = _i32(prob.nx), _i32(prob.nu)
        nc = _i32(prob.nc) if prob.nc is not None else None
        ncp = _ip(nc) if nc is not None else None
        size = L.tree_qp_out_calculate_size(prob.Nn, _ip(nx), _ip(nu), ncp)
        self.mem = _Buf(size)
        self.struct_buf = _Buf(C.sizeof(QpOutStruct) + 64)
        L.tree_qp_out_create(prob.Nn, _ip(nx), _ip(nu), ncp, self.struct_buf.ptr, self.mem.ptr)
        self.ptr = self.struct_buf.ptr
        self.s = QpOutStruct.from_address(self.struct_buf.addr)

    def pin(self) -> bool:
        """page-lock the tree_qp_out memory: large trees then download by DMA straight into it"""
        return self.mem.pin(self.lib)

    def _gather(self, field_ptr, count) -> np.ndarray:
        """Concatenate `count` blasfeo_dvec views starting at field_ptr (all live in self.mem)."""
        if count == 0:
            return np.zeros(0)
        addr = C.cast(field_ptr, C.c_void_p).value
        recs = np.frombuffer((C.c_char * (count * _DVEC_DT.itemsize)).from_address(addr), dtype=_DVEC_DT)
        m = recs["m"].astype(np.int64)
        base = self.mem.addr
        start = (recs["pa"].astype(np.int64) - base) // 8
        assert (start >= 0).all() and ((start + m) * 8 <= self.mem.nbytes).all()
        data = self.mem.arr[: self.mem.nbytes // 8 * 8].view(np.float64)
        tot = int(m.sum())
        if tot == 0:
            return np.zeros(0)
        seg = np.repeat(start - np.concatenate([[0], np.cumsum(m)[:-1]]), m)
        return data[seg + np.arange(tot)].copy()

    @property
    def x(self): return self._gather(self.s.x, self.prob.Nn)
    @property
    def u(self): return self._gather(self.s.u, self.prob.Nn)
    @property
    def lam(self): return self._gather(self.s.lam, self.prob.Nn - 1)
    @property
    def mu_x(self): return self._gather(self.s.mu_x, self.prob.Nn)
    @property
    def mu_u(self): return self._gather(self.s.mu_u, self.prob.Nn)
    @property
    def mu_d(self): return self._gather(self.s.mu_d, self.prob.Nn)
    @property
    def iter(self): return int(self.s.info.iter)
    @property
    def total_time(self): return float(self.s.info.total_time)


@dataclass
class Solution:
    status: int
    iter: int
    x: np.ndarray
    u: np.ndarray
    lam: np.ndarray
    mu_x: np.ndarray
    mu_u: np.ndarray
    kkt: float
    total_time: float = 0.0
    stats: Optional[dict] = None


class Tdunes:
    """TDUNES solver object: opts + workspace protocol of dual_Newton_tree.h:154-168."""

    def __init__(self, lib: TreeQpLib, qp_in: QpIn, **opt_overrides):
        self.lib, self.qp_in = lib, qp_in
        L = lib.lib
        Nn = qp_in.prob.Nn
        self.opts = TdunesOpts()
        self.opts_mem = _Buf(max(L.treeqp_tdunes_opts_calculate_size(Nn), 8))
        L.treeqp_tdunes_opts_create(Nn, C.byref(self.opts), self.opts_mem.ptr)
        L.treeqp_tdunes_opts_set_default(Nn, C.byref(self.opts))
        qp_solver = opt_overrides.pop("qp_solver", None)
        if qp_solver is not None:
            for k in range(Nn):
                self.opts.qp_solver[k] = int(qp_solver if np.isscalar(qp_solver) else qp_solver[k])
        for k, v in opt_overrides.items():
            assert hasattr(self.opts, k), k
            setattr(self.opts, k, v)
        size = L.treeqp_tdunes_calculate_size(qp_in.ptr, C.byref(self.opts))
        assert size > 0, "workspace size overflows int"
        self.work_mem = _Buf(size)
        self.work = _Buf(2048)  # opaque treeqp_tdunes_workspace
        L.treeqp_tdunes_create(qp_in.ptr, C.byref(self.opts), self.work.ptr, self.work_mem.ptr)
        self._alive = True

    def set_dual_initialization(self, lam0: np.ndarray):
        lam0 = _f64(lam0)
        assert lam0.size == self.qp_in.prob.n_lambda
        self._lam0 = lam0
        self.lib.lib.treeqp_tdunes_set_dual_initialization(_dp(lam0), self.work.ptr)

    def solve(self, qp_out: QpOut) -> int:
        return int(self.lib.lib.treeqp_tdunes_solve(self.qp_in.ptr, qp_out.ptr, C.byref(self.opts), self.work.ptr))

    # additive protocol (product library only)
    def upload(self) -> int:
        return int(self.lib.lib.treeqp_tdunes_upload(self.qp_in.ptr, C.byref(self.opts), self.work.ptr))

    def solve_resident(self) -> int:
        return int(self.lib.lib.treeqp_tdunes_solve_resident(C.byref(self.opts), self.work.ptr))

    def download(self, qp_out: QpOut) -> int:
        return int(self.lib.lib.treeqp_tdunes_download(qp_out.ptr, self.work.ptr))

    def update_x0(self) -> int:
        """after qp_in.set_x0(...): re-send only the root and its children"""
        return int(self.lib.lib.treeqp_tdunes_update_x0(self.qp_in.ptr, self.work.ptr))

    def kkt_resident(self) -> float:
        """tree_qp_out_max_KKT_res of the last solve, evaluated on the device"""
        out = np.zeros(1)
        assert self.lib.lib.treeqp_tdunes_kkt_resident(self.work.ptr, _dp(out)) == TREEQP_OK
        return float(out[0])

    def warm_start_resident(self) -> int:
        """dual start of the next solve = multipliers of the last one, without leaving HBM"""
        return int(self.lib.lib.treeqp_tdunes_warm_start_resident(self.work.ptr))

    def stats(self) -> dict:
        st = Stats()
        self.lib.lib.treeqp_tdunes_get_stats(self.work.ptr, C.byref(st))
        return stats_dict(self.lib, st)

    def destroy(self):
        if self._alive and self.lib.is_product:
            self.lib.lib.treeqp_tdunes_destroy(self.work.ptr)
        self._alive = False

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def solve_tdunes(lib: TreeQpLib, prob: Problem, lam0: Optional[np.ndarray] = None, **opts) -> Solution:
    """Build, solve once, read back: the sequence of examples/spring_mass_dual_newton_tree.c:52-158."""
    qp_in = QpIn(lib, prob)
    qp_out = QpOut(lib, prob)
    solver = Tdunes(lib, qp_in, **opts)
    if lam0 is None:
        lam0 = prob.lam0 if prob.lam0 is not None else np.zeros(prob.n_lambda)
    solver.set_dual_initialization(lam0)
    status = solver.solve(qp_out)
    kkt = float("nan")
    if status in (TREEQP_OPTIMAL_SOLUTION_FOUND, TREEQP_MAXIMUM_ITERATIONS_REACHED):
        kkt = float(lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
    sol = Solution(status=status, iter=qp_out.iter, x=qp_out.x, u=qp_out.u, lam=qp_out.lam,
                   mu_x=qp_out.mu_x, mu_u=qp_out.mu_u, kkt=kkt, total_time=qp_out.total_time,
                   stats=solver.stats() if lib.is_product else None)
    solver.destroy()
    return sol


class TdunesBatch:
    """Batch of independent tree QPs solved as one forest (additive API of include/treeqp_b200.h:
    treeqp_tdunes_batch_*): every kernel launch covers the same tree level of all QPs."""

    def __init__(self, lib: TreeQpLib, qp_ins, **opt_overrides):
        assert lib.is_product
        self.lib, self.qp_ins, self.n = lib, list(qp_ins), len(qp_ins)
        L = lib.lib
        Nn = max(q.prob.Nn for q in self.qp_ins)
        self.opts = TdunesOpts()
        self.opts_mem = _Buf(max(L.treeqp_tdunes_opts_calculate_size(Nn), 8))
        L.treeqp_tdunes_opts_create(Nn, C.byref(self.opts), self.opts_mem.ptr)
        L.treeqp_tdunes_opts_set_default(Nn, C.byref(self.opts))
        qp_solver = opt_overrides.pop("qp_solver", None)
        if qp_solver is not None:
            for k in range(Nn):
                self.opts.qp_solver[k] = int(qp_solver if np.isscalar(qp_solver) else qp_solver[k])
        for k, v in opt_overrides.items():
            assert hasattr(self.opts, k), k
            setattr(self.opts, k, v)
        self.in_ptrs = (C.c_void_p * self.n)(*[q.ptr.value for q in self.qp_ins])
        self.handle = C.c_void_p(L.treeqp_tdunes_batch_create(self.n, self.in_ptrs, C.byref(self.opts)))
        if not self.handle.value:
            raise RuntimeError("treeqp_tdunes_batch_create failed (no CUDA device?)")
        self.status = np.zeros(self.n, dtype=np.int32)

    def set_dual_initialization(self, i: int, lam0: np.ndarray):
        lam0 = _f64(lam0)
        self.lib.lib.treeqp_tdunes_batch_set_dual_initialization(self.handle, i, _dp(lam0))

    def upload(self) -> int:
        return int(self.lib.lib.treeqp_tdunes_batch_upload(self.handle, self.in_ptrs))

    def update_x0(self) -> int:
        """after QpIn.set_x0 on every member: re-send only what depends on x0"""
        return int(self.lib.lib.treeqp_tdunes_batch_update_x0(self.handle, self.in_ptrs))

    def solve_resident(self) -> int:
        return int(self.lib.lib.treeqp_tdunes_batch_solve_resident(self.handle, _ip(self.status)))

    def kkt_resident(self) -> np.ndarray:
        """per-QP tree_qp_out_max_KKT_res of the last solve, evaluated on the device"""
        out = np.zeros(self.n)
        assert self.lib.lib.treeqp_tdunes_batch_kkt_resident(self.handle, _dp(out)) == TREEQP_OK
        return out

    def _out_ptrs(self, qp_outs):
        return (C.c_void_p * self.n)(*[q.ptr.value for q in qp_outs])

    def download(self, qp_outs) -> int:
        return int(self.lib.lib.treeqp_tdunes_batch_download(self.handle, self._out_ptrs(qp_outs)))

    def solve(self, qp_outs) -> int:
        """treeqp_tdunes_solve_batch: host buffers in, host buffers out"""
        return int(self.lib.lib.treeqp_tdunes_solve_batch(self.handle, self.in_ptrs, self._out_ptrs(qp_outs),
                                                          _ip(self.status)))

    def stats(self) -> dict:
        st = Stats()
        self.lib.lib.treeqp_tdunes_batch_get_stats(self.handle, C.byref(st))
        return stats_dict(self.lib, st)

    def destroy(self):
        if self.handle and self.handle.value:
            self.lib.lib.treeqp_tdunes_batch_destroy(self.handle)
            self.handle = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def shard_plan(lib: TreeQpLib, qp_in: "QpIn", rank: int, nranks: int) -> dict:
    """treeqp_b200_shard_plan: which nodes rank `rank` of `nranks` owns / sees as ghosts (host logic only)"""
    N = qp_in.prob.Nn
    owned, ghost = np.zeros(N, dtype=np.int32), np.zeros(N, dtype=np.int32)
    cd, nc, nl = C.c_int(), C.c_int(), C.c_int()
    rc = lib.lib.treeqp_b200_shard_plan(qp_in.ptr, rank, nranks, C.byref(cd), C.byref(nc), C.byref(nl), _ip(owned), _ip(ghost))
    if rc != 0:
        raise ValueError("tree cannot be cut for %d ranks" % nranks)
    return dict(cut_depth=cd.value, ncut=nc.value, n_local=nl.value, owned=owned.astype(bool), ghost=ghost.astype(bool))


class Comm:
    """NCCL communicator of the sharded solves; `bcast(bytes_or_None) -> bytes` distributes rank 0's
    128-byte id (e.g. over torch.distributed)."""

    def __init__(self, lib: TreeQpLib, rank: int, nranks: int, device: int, bcast):
        self.lib = lib
        buf = C.create_string_buffer(128)
        if rank == 0:
            assert lib.lib.treeqp_b200_comm_unique_id(buf) == 0, "NCCL unavailable"
        ident = bcast(buf.raw if rank == 0 else None)
        assert len(ident) == 128
        self.handle = C.c_void_p(lib.lib.treeqp_b200_comm_create(rank, nranks, ident, device))
        if not self.handle.value:
            raise RuntimeError("treeqp_b200_comm_create failed")
        self.rank, self.nranks = rank, nranks

    def destroy(self):
        if self.handle and self.handle.value:
            self.lib.lib.treeqp_b200_comm_destroy(self.handle)
            self.handle = C.c_void_p(None)


class TdunesSharded:
    """One tree QP sharded by subtree over the ranks of `comm` (treeqp_tdunes_sharded_*)."""

    def __init__(self, lib: TreeQpLib, qp_in: "QpIn", comm: Comm, **opt_overrides):
        self.lib, self.qp_in, self.comm = lib, qp_in, comm
        L = lib.lib
        Nn = qp_in.prob.Nn
        self.opts = TdunesOpts()
        self.opts_mem = _Buf(max(L.treeqp_tdunes_opts_calculate_size(Nn), 8))
        L.treeqp_tdunes_opts_create(Nn, C.byref(self.opts), self.opts_mem.ptr)
        L.treeqp_tdunes_opts_set_default(Nn, C.byref(self.opts))
        for k, v in opt_overrides.items():
            assert hasattr(self.opts, k), k
            setattr(self.opts, k, v)
        self.handle = C.c_void_p(L.treeqp_tdunes_sharded_create(qp_in.ptr, C.byref(self.opts), comm.handle))
        if not self.handle.value:
            raise RuntimeError("treeqp_tdunes_sharded_create failed")

    def plan(self) -> dict:
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.lib.lib.treeqp_tdunes_sharded_get_plan(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return dict(cut_depth=a.value, ncut=b.value, n_local=c.value, n_owned=d.value)

    def set_dual_initialization(self, lam0: np.ndarray):
        lam0 = _f64(lam0)
        self.lib.lib.treeqp_tdunes_sharded_set_dual_initialization(self.handle, _dp(lam0), self.qp_in.ptr)

    def upload(self) -> int:
        return int(self.lib.lib.treeqp_tdunes_sharded_upload(self.handle, self.qp_in.ptr))

    def solve_resident(self) -> int:
        return int(self.lib.lib.treeqp_tdunes_sharded_solve_resident(self.handle))

    def download(self, qp_out: "QpOut") -> int:
        return int(self.lib.lib.treeqp_tdunes_sharded_download(self.handle, qp_out.ptr))

    def solve(self, qp_out: "QpOut") -> int:
        return int(self.lib.lib.treeqp_tdunes_sharded_solve(self.handle, self.qp_in.ptr, qp_out.ptr))

    def stats(self) -> dict:
        st = Stats()
        self.lib.lib.treeqp_tdunes_sharded_get_stats(self.handle, C.byref(st))
        return stats_dict(self.lib, st)

    def destroy(self):
        if self.handle and self.handle.value:
            self.lib.lib.treeqp_tdunes_sharded_destroy(self.handle)
            self.handle = C.c_void_p(None)


class Sdunes:
    """SDUNES solver object: opts + workspace protocol of dual_Newton_scenarios.h:143-163.
    The QP must have x0 eliminated (nx[0] == 0), uniform nx/nu and a multistage tree."""

    def __init__(self, lib: TreeQpLib, qp_in: QpIn, **opt_overrides):
        self.lib, self.qp_in = lib, qp_in
        L = lib.lib
        prob = qp_in.prob
        Nn = prob.Nn
        self.opts = SdunesOpts()
        self.opts_mem = _Buf(max(L.treeqp_sdunes_opts_calculate_size(Nn), 8))
        L.treeqp_sdunes_opts_create(Nn, C.byref(self.opts), self.opts_mem.ptr)
        L.treeqp_sdunes_opts_set_default(Nn, C.byref(self.opts))
        for k, v in opt_overrides.items():
            assert hasattr(self.opts, k), k
            setattr(self.opts, k, v)
        size = L.treeqp_sdunes_calculate_size(qp_in.ptr, C.byref(self.opts))
        assert size > 0, "workspace size overflows int"
        self.work_mem = _Buf(size)
        self.work = _Buf(2048)  # opaque treeqp_sdunes_workspace
        L.treeqp_sdunes_create(qp_in.ptr, C.byref(self.opts), self.work.ptr, self.work_mem.ptr)
        dims = np.frombuffer((C.c_char * 16).from_address(self.work.addr), dtype=np.int32)
        self.Ns, self.Nh, self.Nr, self.md = (int(v) for v in dims)
        self.nx, self.nu = int(prob.nx[1]), int(prob.nu[0])
        self.n_lambda = max(L.treeqp_sdunes_calculate_dual_dimension(self.Nr, self.md, self.nu), 0)
        self.n_mu = self.Ns * self.Nh * self.nx
        self._alive = True

    def set_dual_initialization(self, lam0: np.ndarray, mu0: np.ndarray):
        lam0, mu0 = _f64(lam0), _f64(mu0)
        assert lam0.size == self.n_lambda and mu0.size == self.n_mu
        self._keep = (lam0, mu0)
        self.lib.lib.treeqp_sdunes_set_dual_initialization(_dp(lam0), _dp(mu0), self.work.ptr)

    def solve(self, qp_out: QpOut) -> int:
        return int(self.lib.lib.treeqp_sdunes_solve(self.qp_in.ptr, qp_out.ptr, C.byref(self.opts), self.work.ptr))

    # additive protocol (product library only)
    def upload(self) -> int:
        return int(self.lib.lib.treeqp_sdunes_upload(self.qp_in.ptr, C.byref(self.opts), self.work.ptr))

    def solve_resident(self) -> int:
        return int(self.lib.lib.treeqp_sdunes_solve_resident(C.byref(self.opts), self.work.ptr))

    def download(self, qp_out: QpOut) -> int:
        return int(self.lib.lib.treeqp_sdunes_download(qp_out.ptr, self.work.ptr))

    def duals(self):
        """scenario duals after the last solve: (lambda [n_lambda], mu [Ns*Nh*nx])"""
        lam, mu = np.zeros(max(self.n_lambda, 1)), np.zeros(max(self.n_mu, 1))
        self.lib.lib.treeqp_sdunes_get_duals(_dp(lam), _dp(mu), self.work.ptr)
        return lam[:self.n_lambda], mu[:self.n_mu]

    def stats(self) -> dict:
        st = Stats()
        self.lib.lib.treeqp_sdunes_get_stats(self.work.ptr, C.byref(st))
        return stats_dict(self.lib, st, sdunes=True)

    def destroy(self):
        if self._alive and self.lib.is_product:
            self.lib.lib.treeqp_sdunes_destroy(self.work.ptr)
        self._alive = False

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def solve_sdunes(lib: TreeQpLib, prob: Problem, lam0=None, mu0=None, **opts) -> Solution:
    """Build, solve once, read back: examples/spring_mass_dual_newton_scenarios.c:100-166."""
    qp_in = QpIn(lib, prob)
    qp_out = QpOut(lib, prob)
    solver = Sdunes(lib, qp_in, **opts)
    solver.set_dual_initialization(np.zeros(solver.n_lambda) if lam0 is None else lam0,
                                   np.zeros(solver.n_mu) if mu0 is None else mu0)
    status = solver.solve(qp_out)
    kkt = float(lib.lib.tree_qp_out_max_KKT_res(qp_in.ptr, qp_out.ptr))
    sol = Solution(status=status, iter=qp_out.iter, x=qp_out.x, u=qp_out.u, lam=qp_out.lam,
                   mu_x=qp_out.mu_x, mu_u=qp_out.mu_u, kkt=kkt, total_time=qp_out.total_time,
                   stats=solver.stats() if lib.is_product else None)
    solver.destroy()
    return sol

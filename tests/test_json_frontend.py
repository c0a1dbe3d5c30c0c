"""tools/solve_qp_json.cpp: the reference's JSON harness format (examples/solve_qp_json.cpp) in front of
libtreeqp_b200.so.  CPU: the file is read and the tree rebuilt; GPU: solutions match the goldens."""
import os

import numpy as np
import pytest

import cases as CS
from treeqp_b200 import problems as P, qp_json as J

needs_exe = pytest.mark.skipif(not os.path.exists(J.EXE), reason="solve_qp_json not built (no nlohmann/json.hpp)")


@needs_exe
def test_reads_harness_format(tmp_path):
    prob = P.spring_mass(3, 2, 10, xmax1=0.2)
    info = J.run_exe(J.problem_to_json(prob, options={"solver": "tdunes", "clipping": True}), str(tmp_path), ["--dry-run"])
    assert info == {"nodes": 85, "solver": "tdunes", "nx0_eliminated": True, "dim_lam_tree": int(prob.n_lambda)}
    # irregular dimensions, scalars and empty arrays as MATLAB's jsonencode writes them
    prob, _, _ = CS.ref_random_qp("01")
    info = J.run_exe(J.problem_to_json(prob), str(tmp_path), ["--dry-run"])
    assert info["nodes"] == prob.Nn and info["nx0_eliminated"] is False and info["dim_lam_tree"] == int(prob.n_lambda)


@needs_exe
@pytest.mark.gpu
def test_tdunes_clipping_through_json(tmp_path):
    prob, _ = CS.get_case("spring_mass_tight")
    g = CS.golden("spring_mass_tight")
    res = J.run_exe(J.problem_to_json(prob, options={"solver": "tdunes", "clipping": True}), str(tmp_path), ["--nrep", "2"])
    assert res["info"]["status"] == 0 and res["info"]["num_iter"] == int(g["iter"]) == 58
    assert res["info"]["kkt_tol"] < 1e-8
    x, u, lam = J.solution_arrays(res)
    assert CS.rel_err(x, g["x"]) < 1e-10 and CS.rel_err(u, g["u"]) < 1e-10 and CS.rel_err(lam, g["lam"]) < 1e-10


@needs_exe
@pytest.mark.gpu
def test_dense_reference_goldens_through_json(tmp_path):
    """the reference's own data0N problems in its own JSON format: xopt/uopt to 1e-12"""
    for n in CS.REF_RANDOM_QP:
        prob, xopt, uopt = CS.ref_random_qp(n)
        opts = {"solver": "tdunes", "clipping": False, "maxit": 10, "stationarityTolerance": 1e-10, "regType": "TREEQP_NO_REGULARIZATION"}
        res = J.run_exe(J.problem_to_json(prob, options=opts), str(tmp_path))
        assert res["info"]["status"] == 0 and res["info"]["num_iter"] <= 1
        x, u, _ = J.solution_arrays(res)
        assert np.max(np.abs(x - xopt)) < 1e-12 and np.max(np.abs(u - uopt)) < 1e-12


@needs_exe
@pytest.mark.gpu
def test_sdunes_through_json(tmp_path):
    prob = P.spring_mass(3, 2, 10)      # pinned x0: the harness eliminates it
    g = CS.sdunes_golden("spring_mass_cfg1")
    res = J.run_exe(J.problem_to_json(prob, options={"solver": "sdunes"}), str(tmp_path))
    assert res["info"]["status"] == 0 and res["info"]["num_iter"] == int(g["iter"]) == 3
    x, u, _ = J.solution_arrays(res)
    assert CS.rel_err(x[4:], g["x"]) < 1e-10 and CS.rel_err(u, g["u"]) < 1e-10
    assert np.array_equal(x[:4], prob.xmin[:4])

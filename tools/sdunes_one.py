import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases as CS
from treeqp_b200 import capi
lib = capi.TreeQpLib(capi.PRODUCT_LIB)
prob, opts = CS.get_sdunes_case("spring_mass_cfg2")
qp_in, qp_out = capi.QpIn(lib, prob), capi.QpOut(lib, prob)
s = capi.Sdunes(lib, qp_in, **opts)
s.set_dual_initialization(np.zeros(s.n_lambda), np.zeros(s.n_mu))
s.upload()
for r in range(2):
    s.solve_resident()
print(s.stats()["gpu_ms"])

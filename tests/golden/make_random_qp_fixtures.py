"""Turns the reference's OWN golden vectors for the dense TDUNES path into small fixtures:
examples/random_qp_utils/data00.c ... data05.c (problem data + xopt/uopt, checked by
examples/random_qp.c:196-254 to 1e-12 with maxIter=10, stationarityTolerance=1e-10, no
regularisation, qp_solver = TREEQP_QPOASES_SOLVER, infinite bounds).

    python tests/golden/make_random_qp_fixtures.py        (needs /root/reference; run in the build container)

Writes tests/golden/refdata_random_qp_0N.npz: the arrays exactly as the .c files list them (the
order tree_qp_in_set_ltv_* consumes).  Nothing is computed here."""
import os
import re

import numpy as np

SRC = "/root/reference/examples/random_qp_utils"
DST = os.path.dirname(os.path.abspath(__file__))

for n in range(6):
    text = open(os.path.join(SRC, "data%02d.c" % n)).read()
    arrays = {}
    for typ, name, body in re.findall(r"(int|double)\s+(\w+)\[\d*\]\s*=\s*\{([^}]*)\}", text):
        vals = [t for t in body.replace("\n", " ").split(",") if t.strip()]
        arrays[name] = np.array([int(t) if typ == "int" else float(t) for t in vals], dtype=np.int32 if typ == "int" else np.float64)
    Nn = int(re.search(r"int\s+Nn\s*=\s*(\d+)", text).group(1))
    assert "UNCONSTRAINED" in text and len(arrays["nk"]) == Nn
    need = ("nk", "nx", "nu", "A", "B", "b", "Q", "R", "S", "q", "r", "xopt", "uopt")
    np.savez(os.path.join(DST, "refdata_random_qp_%02d.npz" % n), **{k: arrays[k] for k in need})
    print(n, Nn, {k: arrays[k].size for k in need})

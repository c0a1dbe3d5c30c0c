#!/bin/bash
# time the depth-16 benchmark QP (resident solves) with every library variant under treeqp_b200/lib
cd "$(dirname "$0")/.."
for lib in treeqp_b200/lib/libtreeqp_b200.so treeqp_b200/lib/var_*/libtreeqp_b200.so; do
  [ -f "$lib" ] || continue
  echo "== $lib"
  TREEQP_B200_LIB=$PWD/$lib python tools/profile_solve.py --depth 16 --solves 3 2>&1 | tail -1
done

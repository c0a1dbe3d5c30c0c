#!/bin/bash
# usage: tools/sass_count.sh <mangled-substring>  -- SASS instruction count and register/spill line of one kernel
O=/root/repo/treeqp_b200/build/device.o
cuobjdump -sass $O 2>/dev/null | awk -v k="$1" '/Function :/{f=index($0,k)>0} f' | grep -c "^ *\/\*[0-9a-f]\{4,6\}\*\/"
grep -A2 "Compiling entry function '[^']*$1" /root/repo/treeqp_b200/build/device.ptxas.log | grep -v Compiling

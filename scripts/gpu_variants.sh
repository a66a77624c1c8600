#!/bin/bash
# device-leg time of the fast lane for kernel build variants (squarna_b200/x_*.so, see csrc/Makefile XFLAGS)
for lib in ${LIBS:-libsqrn_b200.so x_RSP.so x_LEV.so x_FIN.so x_ALL.so}; do
  SQRN_LIB_PATH=$PWD/squarna_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-cli 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); e=l['e2e']; print('$lib: device %.3f ms  e2e packed %.3f ms (kernels %.3f)  bytes %.3f ms' % (l['ms_per_step'], e['ms_per_step'], e['kernel_ms_in_step'], e['byte_format']['ms_per_step']))"
done

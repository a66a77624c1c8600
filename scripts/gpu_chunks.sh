#!/bin/bash
# end-to-end leg against the number of pipeline chunks (SQRN_FAST_CHUNKS)
for c in ${CHUNKS:-4 8 12 18}; do
  echo "== chunks $c"
  SQRN_FAST_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-cli 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); e=l['e2e']; print('device %.3f ms  packed %.3f ms (kernels %.3f)  bytes %.3f ms' % (l['ms_per_step'], e['ms_per_step'], e['kernel_ms_in_step'], e['byte_format']['ms_per_step']))"
done

#!/bin/bash
for i in 1 2; do
timeout 600 python bench.py --config 3 --seqs ${C3_SEQS:-2000} --steps 1 --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('value %.1f seq/s (%.1f ms)  e2e %.1f seq/s (%.1f ms)  kernel_ms %.1f' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['kernel_ms']))"
done
python - <<'PY'
import time, sys, cProfile, pstats
import numpy as np
import workloads, bench
from squarna_b200 import SQRNdbnseq as S
entries = workloads.config3(2000, bench.SEED)
groups = {}
for e in entries:
    groups.setdefault(workloads.config3_conf(len(e[0])), []).append((e[0], e[1], e[2], None))
psets = {c: bench.conf_gsets(c) for c in groups}
for c, es in groups.items():
    S.predict_many(es[:8], psets[c], poollim=100, device=0)
for rep in range(2):
    t=time.perf_counter()
    for c, es in groups.items():
        t1=time.perf_counter(); S.predict_many(es, psets[c], poollim=100, device=0); print(c, len(es), "%.3f s" % (time.perf_counter()-t1))
    print("total %.3f" % (time.perf_counter()-t))
pr=cProfile.Profile(); pr.enable()
for c, es in groups.items():
    S.predict_many(es, psets[c], poollim=100, device=0)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
PY

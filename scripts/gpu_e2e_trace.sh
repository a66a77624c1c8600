#!/bin/bash
mkdir -p gpurun_out
for c in 6 8 10 12 15 18; do echo "chunks=$c"; SQRN_FAST_CHUNKS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])"; done 2>&1 | tee gpurun_out/e2e_chunks.log

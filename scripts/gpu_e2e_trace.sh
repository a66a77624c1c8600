#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
SQRN_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-1500; tail -22 gpurun_out/bench.err
for c in 4 8 16; do echo "chunks=$c"; SQRN_FAST_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])"; done 2>&1 | tee gpurun_out/e2e_chunks.log

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for c in 1 2 4 8 16; do echo "chunks=$c"; SQRN_FAST_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value ms', d['ms_per_step'], 'e2e', d['e2e'])"; done 2>&1 | tee gpurun_out/e2e_chunks.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 3 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --seqs 100000 --no-cpu > gpurun_out/ncu_full.log 2>&1

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --seqs 100000 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_work -s 1 -c 1 -f -o gpurun_out/prof_long \
    python scripts/bench_configs.py c5 --n 148 --lo 2900 --hi 3100 > gpurun_out/ncu_long.log 2>&1
tail -2 gpurun_out/ncu_long.log

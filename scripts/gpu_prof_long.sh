#!/bin/bash
# ncu capture of k_long<32> on 148 rRNA-scale sequences (one CTA each)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_long -c 1 -f -o gpurun_out/prof_long \
    python scripts/bench_configs.py c5 --n 148 > gpurun_out/ncu_long.log 2>&1
tail -2 gpurun_out/ncu_long.log

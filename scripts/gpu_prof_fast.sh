#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_fast$ -s 3 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --seqs 200000 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log

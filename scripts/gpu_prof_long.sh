#!/bin/bash
# ncu --set full capture of k_long<32> on 148 rRNA-scale sequences (one CTA each) through bench.py --config 5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_long$ -s 3 -c 1 -f -o gpurun_out/prof_long \
    python bench.py --config 5 --seqs 148 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_long.log 2>&1
tail -3 gpurun_out/ncu_long.log

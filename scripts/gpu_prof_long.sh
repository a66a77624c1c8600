#!/bin/bash
# ncu capture of k_long<32> on 148 rRNA-scale sequences (one CTA each) + the reduced-size config legs
mkdir -p gpurun_out
{
timeout 600 python scripts/bench_configs.py mid --n 4000
SQRN_TRACE=1 timeout 600 python scripts/bench_configs.py c5 --n 592 2>&1 | grep -v "chunk\|fast_predict_host"
timeout 900 python scripts/bench_configs.py c3 --n 96
timeout 900 python scripts/bench_configs.py c4 --n 400
} 2>&1 | tee gpurun_out/configs.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_long -s 1 -c 1 -f -o gpurun_out/prof_long \
    python scripts/bench_configs.py c5 --n 148 > gpurun_out/ncu_long.log 2>&1
tail -2 gpurun_out/ncu_long.log

#!/bin/bash
# launch list (per-kernel durations) of the config-3 bench on a reduced batch
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/c3_launches.csv \
    python bench.py --config 3 --seqs ${C3_SEQS:-600} --steps 1 --warmup 0 --no-cpu > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log | cut -c1-600
timeout 600 python bench.py --config 3 --seqs ${C3_SEQS2:-2000} --steps 1 --no-cpu 2> gpurun_out/c3.err | cut -c1-900
tail -3 gpurun_out/c3.err

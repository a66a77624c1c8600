#!/bin/bash
# quick GPU check: parity tests + headline bench (no CPU baseline leg) + source-level profile of k_fast
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_fast$ -s 3 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 3 --seqs 200000 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log

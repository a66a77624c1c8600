#!/bin/bash
# GPU-box visit (round 1, session d): parity, smoke, headline bench, the other configs at reduced size.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
{
timeout 600 python scripts/bench_configs.py mid --n 4000
timeout 600 python scripts/bench_configs.py c5 --n 148
timeout 600 python scripts/bench_configs.py c5 --n 8 --oracle 2 --lo 2900 --hi 3000
timeout 900 python scripts/bench_configs.py c3 --n 96
timeout 900 python scripts/bench_configs.py c4 --n 400
} 2>&1 | tee gpurun_out/configs.log

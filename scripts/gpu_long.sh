#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
{
timeout 600 python scripts/bench_configs.py mid --n 4000
timeout 600 python scripts/bench_configs.py c5 --n 148
timeout 900 python scripts/bench_configs.py c3 --n 96
} 2>&1 | tee gpurun_out/configs.log

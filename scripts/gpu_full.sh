#!/bin/bash
# full -m gpu suite, smoke, the three bench configs (no CPU baselines)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -16 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-cli 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-700
timeout 900 python bench.py --config 5 --seqs ${C5_SEQS:-592} --steps 1 --warmup 3 --no-cpu 2> gpurun_out/bench_c5.err | tee gpurun_out/bench_c5.json | cut -c1-400
timeout 900 python bench.py --config 3 --seqs ${C3_SEQS:-2000} --steps 1 --no-cpu 2> gpurun_out/bench_c3.err | tee gpurun_out/bench_c3.json | cut -c1-400
tail -3 gpurun_out/bench.err gpurun_out/bench_c5.err gpurun_out/bench_c3.err

#!/bin/bash
# packed boundary format: parity tests, then the config-2 bench (byte and packed end-to-end legs)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "packed or fast_lane_short or fast_kernel_equals" --durations=5 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_FLAGS:---no-cpu --no-cli} > gpurun_out/bench_packed.json 2> gpurun_out/bench_packed.err; cat gpurun_out/bench_packed.json; tail -5 gpurun_out/bench_packed.err

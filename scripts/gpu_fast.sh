#!/bin/bash
# fast lane: parity tests, then the config-2 bench line (no CPU legs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "fast_lane or fast_kernel or packed or config2" --durations=4 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-cli 2> gpurun_out/bench_fast.err | tee gpurun_out/bench_fast.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); e=l['e2e']; print('device %.3f ms (%.1f M seq/s)  e2e packed %.3f ms (%.1f M seq/s, kernels %.3f)  bytes %.3f ms' % (l['ms_per_step'], l['value']/1e6, e['ms_per_step'], e['value']/1e6, e['kernel_ms_in_step'], e['byte_format']['ms_per_step']))"
tail -3 gpurun_out/bench_fast.err

#!/bin/bash
# closing check of the round: the whole -m gpu suite, smoke, both arms of the default bench
T=${TAG:-r02d}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=6 2>&1 | tail -12 > gpurun_out/${T}_pytest_gpu.log; tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 > gpurun_out/${T}_smoke.log; cut -c1-200 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 900 python bench.py --config 5 --steps 1 --warmup 3 --no-cpu > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err; cut -c1-300 gpurun_out/${T}_bench_c5.json; tail -3 gpurun_out/${T}_bench_c5.err

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "cluster or global_list" 2>&1 | tail -4
for n in 1 4 16; do
  timeout 300 python scripts/bench_configs.py c5 --n $n --lo 4000 --hi 5000 2>&1 | tail -1 | cut -c1-260
  SQRN_CLUSTER=1 timeout 300 python scripts/bench_configs.py c5 --n $n --lo 4000 --hi 5000 2>&1 | tail -1 | cut -c1-260
done

#!/bin/bash
# pool rounds (config 3): parity tests, then the config-3 bench line with and without the base lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "${PYTEST_K:-predict_batch_full or config3 or long_reference or bpp or non_greedy or entropy or optimal_step}" --durations=6 2>&1 | tail -12
for nb in ${NOBASE:-0 1}; do
  echo "== SQRN_NO_BASE=$nb"
  if [ "$nb" = 1 ]; then export SQRN_NO_BASE=1; else unset SQRN_NO_BASE; fi
  timeout 600 python bench.py --config 3 --seqs ${C3_SEQS:-2000} --steps 1 --no-cpu 2> gpurun_out/c3_$nb.err | tee gpurun_out/c3_$nb.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('value %.1f seq/s (%.1f ms)  e2e %.1f seq/s (%.1f ms)  kernel_ms %.1f  launches %s  calls %s' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['kernel_ms'], l['gpu_launches'], l['roofline']['optimal_calls_per_step']))"
  tail -2 gpurun_out/c3_$nb.err
done

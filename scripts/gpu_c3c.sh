#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "predict_batch_full or config3 or long_reference or bpp or non_greedy or entropy or optimal_step or cli" --durations=4 2>&1 | tail -8
SQRN_TRACE=1 timeout 600 python bench.py --config 3 --seqs ${C3_SEQS:-2000} --steps 1 --no-cpu 2> gpurun_out/c3_t.err | tee gpurun_out/c3_t.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('value %.1f seq/s (%.1f ms)  e2e %.1f seq/s (%.1f ms)  kernel_ms %.1f  launches %s  calls %s' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['kernel_ms'], l['gpu_launches'], l['roofline']['optimal_calls_per_step']))"
grep "predict_batch" gpurun_out/c3_t.err | tail -3

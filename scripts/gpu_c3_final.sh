#!/bin/bash
# config 3: pool parity tests, the trace of the C call, then the bench line with the bounded CPU baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "predict_batch_full or config3 or long_reference or bpp or non_greedy or cli" 2>&1 | tail -2
SQRN_TRACE=1 timeout 600 python bench.py --config 3 --seqs 2000 --steps 1 --no-cpu 2>gpurun_out/c3v.err | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('value %.1f seq/s (%.1f ms)  e2e %.1f seq/s (%.1f ms)  kernel_ms %.1f' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['roofline']['kernel_ms']))"
grep "800 sequences" gpurun_out/c3v.err | tail -1 | cut -c1-330
timeout 1100 python bench.py --config 3 --seqs 4000 --steps 1 > gpurun_out/r02c_bench_c3.json 2> gpurun_out/r02c_bench_c3.err
python -c "
import json; l=json.load(open('gpurun_out/r02c_bench_c3.json')); print(l['value'], l['e2e']['value'], l['roofline']['kernel_ms'], l['cpu_baseline'])"; tail -3 gpurun_out/r02c_bench_c3.err

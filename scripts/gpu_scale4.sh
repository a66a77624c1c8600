#!/bin/bash
# the driver's multi-GPU launch at N = 4 (config 2, both arms' rank handling)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/scale4.err | tee gpurun_out/scale4.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); e=l['e2e']; print('N=4: device %.1f M seq/s (%.3f ms)  e2e %.1f M seq/s (%.3f ms)  bytes %.1f M seq/s; bound %s' % (l['value']/1e6, l['ms_per_step'], e['value']/1e6, e['ms_per_step'], e['byte_format']['value']/1e6, l['config']['host_cpus_bound_to_gpu_locality']))"
tail -2 gpurun_out/scale4.err

#!/bin/bash
# the driver's multi-GPU launch at N = 2 (and the reference arm's rank handling)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/scale2.err | tee gpurun_out/scale2.json | cut -c1-1200
tail -3 gpurun_out/scale2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --config 5 --seqs 1184 --gpus 2 --steps 1 --warmup 3 2> gpurun_out/scale2_c5.err | tee gpurun_out/scale2_c5.json | cut -c1-700
tail -3 gpurun_out/scale2_c5.err
timeout 600 python -m pytest tests -m gpu -q -x -k "multigpu or sharding or devices" 2>&1 | tail -4

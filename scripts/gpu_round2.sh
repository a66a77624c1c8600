#!/bin/bash
# Round-2 GPU-box visit: parity tests, smoke, the three bench configs, launch list, ncu capture of k_long.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q -x ${PYTEST_K:+-k "$PYTEST_K"} --durations=12 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 python bench.py --config 5 --seqs ${C5_SEQS:-592} --steps 1 --warmup 3 ${C5_FLAGS:---no-cpu} > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cat gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
timeout 900 python bench.py --config 3 --seqs ${C3_SEQS:-2000} --steps 1 ${C3_FLAGS:---no-cpu} > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --seqs 100000 --no-cpu --no-cli > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_long$ -s 3 -c 1 -f -o gpurun_out/prof_long \
    python bench.py --config 5 --seqs 148 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_long.log 2>&1
tail -3 gpurun_out/ncu_long.log
fi
ls -la gpurun_out

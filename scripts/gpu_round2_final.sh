#!/bin/bash
# Round-2 final GPU-box visit: parity tests, smoke, every bench config (with the CPU baselines), the reference arm,
# launch list and ncu --set full captures of k_fast (packed), k_long<32> and a k_work<8> pool round.
T=${TAG:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_gpu.txt 2>&1
nproc >> gpurun_out/${T}_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${T}_gpu.txt
timeout 2400 python -m pytest tests -m gpu -q -x --durations=10 2>&1 | tail -18 > gpurun_out/${T}_pytest_gpu.log; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 > gpurun_out/${T}_smoke.log; cut -c1-300 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
timeout 900 python bench.py --config 5 --steps 1 --warmup 3 > gpurun_out/${T}_bench_c5.json 2> gpurun_out/${T}_bench_c5.err; cut -c1-400 gpurun_out/${T}_bench_c5.json; tail -3 gpurun_out/${T}_bench_c5.err
timeout 900 python bench.py --config 3 --seqs ${C3_SEQS:-4000} --steps 1 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err; cut -c1-400 gpurun_out/${T}_bench_c3.json; tail -3 gpurun_out/${T}_bench_c3.err
timeout 600 python bench.py --config 4 --steps 2 > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err; cut -c1-400 gpurun_out/${T}_bench_c4.json; tail -3 gpurun_out/${T}_bench_c4.err
timeout 600 python scripts/bench_cli.py 1000000 5000 > gpurun_out/${T}_bench_cli.json 2>/dev/null; cat gpurun_out/${T}_bench_cli.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches_fastlane_100k.csv \
    python bench.py --steps 2 --warmup 3 --seqs 100000 --no-cpu --no-cli > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_fast$ -s 4 -c 1 -f -o gpurun_out/prof_fast_packed \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-cli > gpurun_out/ncu_fast.log 2>&1; tail -2 gpurun_out/ncu_fast.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_long$ -s 3 -c 1 -f -o gpurun_out/prof_long \
    python bench.py --config 5 --seqs 148 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_long.log 2>&1; tail -2 gpurun_out/ncu_long.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_work$ -s ${C3_SKIP:-150} -c 1 -f -o gpurun_out/prof_pool \
    python bench.py --config 3 --seqs 600 --steps 1 --no-cpu > gpurun_out/ncu_pool.log 2>&1; tail -2 gpurun_out/ncu_pool.log
fi
ls -la gpurun_out | tail -30

#!/bin/bash
mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i -E "icc|icache|inst_cache|l1i|gcc|instruction_cache|l0i|__inst_fetch|ifetch" > gpurun_out/icache_metrics.txt
wc -l gpurun_out/icache_metrics.txt
head -60 gpurun_out/icache_metrics.txt

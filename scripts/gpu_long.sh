#!/bin/bash
# k_long iteration visit: the long-sequence parity tests, then config 5 with the list statistics at several rebuild periods
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "${PYTEST_K:-long or config5 or cta_teams or rrna or cluster or config3}" --durations=8 2>&1 | tail -20 > gpurun_out/pytest_long.log
tail -12 gpurun_out/pytest_long.log
for R in ${REBUILDS:-0 8 32}; do
  echo "== rebuild period $R"
  SQRN_GL_REBUILD=$R SQRN_TRACE=1 timeout 600 python bench.py --config 5 --seqs ${C5_SEQS:-592} --steps 1 --warmup 1 --no-cpu 2> gpurun_out/c5_R$R.err | tee gpurun_out/c5_R$R.json | cut -c1-330
  grep -h "sqrn" gpurun_out/c5_R$R.err | tail -4
done

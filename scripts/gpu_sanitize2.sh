#!/bin/bash
mkdir -p gpurun_out
sed -n '/^cat > \/tmp\/san.py/,/^PY$/p' scripts/gpu_sanitize.sh | sed '1d;$d' > /tmp/san.py
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 python /tmp/san.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -2

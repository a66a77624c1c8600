#!/bin/bash
# alignment mode: step 1 on the device (parity with the host accumulation, CLI goldens), timing of config 4
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "alignment_step1 or config4 or ali or yield" --durations=6 2>&1 | tail -12
python - <<'PY'
import time, io, sys, os
import workloads
from squarna_b200 import SQRNdbnali as A, SQUARNA as CLI
rows, ref = workloads.config4(2000, 300, 400)
path = "/tmp/c4.afa"
open(path, "w").write(workloads.config4_text(rows, ref))
for env in ("", "1"):
    if env: os.environ["SQRN_HOST_STEMMATRIX"] = env
    for rep in range(2):
        buf = io.StringIO(); t = time.perf_counter()
        CLI.Predict(inputfile=path, alignment=True, write_to=buf, step3="1")
        t1 = time.perf_counter() - t
    print("config 4 (2000 x 400), step 1 only, %s accumulation: %.3f s" % ("host" if env else "device", t1))
    out1 = buf.getvalue()
    if env: assert out1 == first
    first = out1
PY

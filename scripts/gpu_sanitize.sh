#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small invocations of every kernel family
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from squarna_b200 import SQRNdbnseq as S
from squarna_b200._abi import pack_sequences
from tests import common as T
ctx = S.get_context(0)
seqs = T.rand_seqs(1, 96, 20, 200) + ["G" * 40 + "AAAA" + "C" * 40 + "AAAA" + "G" * 40]
sym, off = pack_sequences(seqs)
a = ctx.fast_predict(T.FASTEST, sym, off)                       # k_fast (+ k_fast_rescan)
b = ctx.fast_predict(T.DEFG1, sym, off)                         # minlen 2: list overflow -> k_fast_rescan
seqs2 = T.rand_seqs(2, 6, 330, 700) + T.rand_seqs(3, 2, 2100, 2300)
sym2, off2 = pack_sequences(seqs2)
c = ctx.fast_predict(T.FASTEST, sym2, off2)                     # k_long<8>, k_long<32>
ctx.set_no_glist(True)
d = ctx.fast_predict(T.FASTEST, sym2, off2)                     # k_work<8>, k_work<32>
ctx.set_no_glist(False)
assert all(np.array_equal(x, y) for x, y in zip(c, d))
import random
rng = random.Random(4)
cases = [T.rand_case(rng, 20, 120) for _ in range(12)]
S.predict_many([(x[0], x[1], x[2], None) for x in cases], [T.DEFG1, T.DEFG2], poollim=20)      # STEP / FINAL / TAIL with init stems
print("sanitize workload ok")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 python /tmp/san.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck.log

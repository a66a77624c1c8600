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
# round 2: packed boundary format (k_fast<.., packed>, k_widen_offsets), pool rounds over base lists with the
# cp.async.bulk / mbarrier ring (k_work<8>: MODE_BASE, STEP, tails), the binned list of k_long with rebuilds, the device
# stem matrix of the alignment mode
from squarna_b200 import _lib, SQRNdbnali as A
packed, _ = _lib.pack_symbols(sym)
e = ctx.fast_predict_packed(T.FASTEST, packed, off.astype(np.uint32))
assert np.array_equal(_lib.unpack_dbn(off.astype(np.uint32), e[0]), a[0])
cases3 = [T.rand_case(rng, 330, 420, p_sep=0.0, p_gap=0.0) for _ in range(3)]
S.predict_many([(x[0], x[1], x[2], None) for x in cases3], [T.G500_1], poollim=12)
ctx.L.sqrn_ctx_set_tuning(ctx.h, 5, 3)                            # rebuild the binned list every 3 passes
f = ctx.fast_predict(T.FASTEST, sym2, off2)
ctx.L.sqrn_ctx_set_tuning(ctx.h, 5, 0)
assert all(np.array_equal(x, y) for x, y in zip(c, f))
import workloads
rows, _ref = workloads.config4(16, 80, 100, seed=9)
ents = [(r_, None, None) for r_ in rows]
m1, cells = A._yield_many(ents, T.ALI["bpweights"], False, T.ALI["minlen"], T.ALI["minbpscore"], device=0, matrix=(len(rows[0]), 20.0))
m2 = A._accumulate_host(A._yield_many(ents, T.ALI["bpweights"], False, T.ALI["minlen"], T.ALI["minbpscore"], device=0), len(rows[0]))
assert (m1 == m2).all()
print("sanitize workload ok")
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 python /tmp/san.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/racecheck.log

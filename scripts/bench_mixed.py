"""Fast lane on a batch of mixed lengths: mostly warp-team sequences with a few longer ones scattered among them
(what a real FASTA looks like).  SQRN_FAST_NO_SPLIT=1 shows the one-kernel-per-chunk behaviour for comparison."""
import random
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from squarna_b200 import _lib
from squarna_b200._abi import pack_sequences
from tests import common as T

n, n_long = int(sys.argv[1]) if len(sys.argv) > 1 else 500000, int(sys.argv[2]) if len(sys.argv) > 2 else 500
rng = random.Random(5)
seqs = T.rand_seqs(1, n, 30, 200)
for k, s in zip(rng.sample(range(n), n_long), T.rand_seqs(2, n_long, 321, 1000)):
    seqs[k] = s
sym, off = pack_sequences(seqs)
ctx = _lib.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    dbn, scores, nst = ctx.fast_predict(T.FASTEST, sym, off)
    dt = time.perf_counter() - t0
    print("mixed %d + %d long: %.1f ms  (%.2f M seq/s)" % (n - n_long, n_long, dt * 1e3, n / dt / 1e6), flush=True)

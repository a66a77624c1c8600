"""bench.py's CLI leg alone (config 2: Predict(inputfile=<1 M-sequence FASTA>, c=fastest, byseq, pl=1), text file in, text
file out), checked against the C-ABI byte lane on the first 2000 sequences.  python scripts/bench_cli_leg.py [n]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                              # noqa: E402
import workloads                                          # noqa: E402
from squarna_b200 import SQRNdbnseq as S                  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
sym, off, lens = workloads.config2(n)
dbn, scores, nst = S.get_context(0).fast_predict(bench.FASTEST, sym, off)
os.environ["SQRN_TRACE"] = "1"
print(json.dumps(bench.cli_leg(sym, off, lens, dbn)))

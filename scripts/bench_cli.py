#!/usr/bin/env python
"""CLI end to end (text file in, text out) on a synthetic FASTA of config-2 sequences:
the bulk text lane against the per-entry path.  python scripts/bench_cli.py [n_bulk] [n_entry]"""
import io, json, os, random, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squarna_b200 import SQUARNA as CLI          # noqa: E402


def make(n, path, seed=20261017):
    rng = random.Random(seed)
    with open(path, "w") as f:
        for k in range(n):
            f.write(">seq%d\n%s\n" % (k, "".join(rng.choice("ACGU") for _ in range(rng.randint(60, 200)))))


def run(path, nobulk):
    if nobulk:
        os.environ["SQRN_NO_BULK"] = "1"
    else:
        os.environ.pop("SQRN_NO_BULK", None)
    with open(os.devnull, "w") as sink:
        t0 = time.perf_counter()
        CLI.Predict(inputfile=path, fileformat="default", configfile="fastest", byseq=True, poollim=1, write_to=sink)
        return time.perf_counter() - t0


n_bulk = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_entry = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
d = tempfile.mkdtemp()
small, big = os.path.join(d, "small.fas"), os.path.join(d, "big.fas")
make(n_entry, small); make(n_bulk, big)
run(small, False)                                  # warm-up: context, module load
t_entry = run(small, True)
t_bulk = min(run(big, False) for _ in range(2))
print(json.dumps({"cli": "i=<fasta> c=fastest byseq pl=1", "bulk_lane": {"n": n_bulk, "seconds": t_bulk, "seq_per_s": n_bulk / t_bulk},
                  "entry_path": {"n": n_entry, "seconds": t_entry, "seq_per_s": n_entry / t_entry}}))

#!/usr/bin/env python
"""Child process of bench.py: runs the UNMODIFIED reference (baseline/_ref/SQUARNA, installed by
scripts/install_reference.py) through its own public API -- Predict(), the function its CLI calls
(SQUARNA.py:416-991) -- on an input file, with the given keyword arguments, and prints one JSON line with the
seconds Predict() took (its multiprocessing.Pool start-up included, the interpreter and numpy imports not).
None of this repository's code is on that path."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(os.path.dirname(HERE), "baseline", "_ref", "SQUARNA")


def main():
    inp, outp, kwargs = sys.argv[1], sys.argv[2], json.loads(sys.argv[3])
    sys.path.insert(0, REF)
    import SQUARNA as REFCLI                      # the reference's SQUARNA.py
    with open(outp, "w") as sink:
        t0 = time.perf_counter()
        REFCLI.Predict(inputfile=inp, write_to=sink, HOME_DIR=REF, **kwargs)
        dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "threads": kwargs.get("threads")}), flush=True)


if __name__ == "__main__":
    main()

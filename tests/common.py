"""Shared helpers of the test-suite: parameter sets, seeded generators."""
import random

import numpy as np


def _ps(**kw):
    base = dict(algorithms={"G"}, bpp=0.0, bpweights={"GC": 3.25, "AU": 1.25, "GU": -1.25},
                suboptmax=0.9, suboptmin=0.65, suboptsteps=1.0, minlen=2.0, minbpscore=4.5,
                minfinscorefactor=1.25, distcoef=0.09, bracketweight=-2.0, orderpenalty=1.0,
                loopbonus=0.125, maxstemnum=1e6)
    base.update(kw)
    return base


# the values of the reference's shipped .conf files (fastest.conf, def.conf G sets, ali.conf, 1000.conf)
FASTEST = _ps(suboptmax=1.0, suboptmin=1.0, minlen=4.0, minbpscore=7.0)
DEFG1 = _ps()
DEFG2 = _ps(bpweights={"GC": 2.0, "AU": 1.0, "GU": 1.0}, minbpscore=3.0, minfinscorefactor=0.99,
            distcoef=0.1, orderpenalty=1.35)
ALI = _ps(bpweights={"GC": 3.25, "AU": 2.0, "GU": -1.0}, suboptmax=1.0, suboptmin=1.0,
          minfinscorefactor=1.0, orderpenalty=0.75)
G1000 = _ps(suboptmax=0.99, suboptmin=0.99)
G500_1 = _ps(suboptmax=0.95, suboptmin=0.9)


def rand_seq(rng, n, alphabet="ACGU"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def rand_seqs(seed, count, lo, hi, alphabet="ACGU"):
    rng = random.Random(seed)
    return [rand_seq(rng, rng.randint(lo, hi), alphabet) for _ in range(count)]


def rand_case(rng, lo, hi, p_sep=0.3, p_rest=0.5, p_react=0.5, p_gap=0.3):
    """A random SQRNdbnseq input exercising separators, restraints, reactivities and gaps."""
    n = rng.randint(lo, hi)
    alpha = "ACGU" if rng.random() < 0.7 else "ACGUTNacgu"
    seq = [rng.choice(alpha) for _ in range(n)]
    if rng.random() < p_sep:
        for _ in range(rng.randint(1, 3)):
            seq[rng.randrange(n)] = rng.choice(";&")
    rest = reacts = None
    if rng.random() < p_rest:
        rest = ["."] * n
        for k in range(n):
            r = rng.random()
            if r < 0.05:
                rest[k] = "_"
            elif r < 0.06:
                rest[k] = "/"
            elif r < 0.07:
                rest[k] = "\\"
            elif r < 0.075:
                rest[k] = "+"
        for _ in range(rng.randint(0, 3)):
            ln, i, j = rng.randint(1, 6), rng.randrange(n), rng.randrange(n)
            if i > j:
                i, j = j, i
            br = rng.choice(["()", "[]", "{}", "<>", "Aa"])
            for k in range(ln):
                if i + k < j - k - 3 and rest[i + k] == "." and rest[j - k] == ".":
                    rest[i + k], rest[j - k] = br[0], br[1]
        rest = "".join(rest)
    if rng.random() < p_react:
        if rng.random() < 0.5:
            reacts = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz?") for _ in range(n))
        else:
            reacts = [round(rng.random(), 3) for _ in range(n)]
    if rng.random() < p_gap:
        for _ in range(rng.randint(1, 5)):
            seq[rng.randrange(n)] = rng.choice("-.~")
    kw = dict(hardrest=rng.random() < 0.3, interchainonly=rng.random() < 0.15, rankbydiff=rng.random() < 0.3,
              conslim=rng.choice([1, 1, 2, 3]), rankby=rng.choice([(0, 2, 1), (1, 0, 2), (2, 1, 0)]))
    return "".join(seq), reacts, rest, kw


def same_prediction(a, b):
    """compare (cons, [(dbn, scores, psinds)]) tuples exactly, incl. the int-0 struct score"""
    if a[0] != b[0] or len(a[1]) != len(b[1]):
        return False
    for x, y in zip(a[1], b[1]):
        if x[0] != y[0] or tuple(x[1]) != tuple(y[1]) or list(x[2]) != list(y[2]):
            return False
        if type(x[1][1]) is not type(y[1][1]):
            return False
    return True

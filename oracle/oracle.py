"""ctypes front-end of the CPU oracle (oracle/sqrn_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  The product package
(squarna_b200) never imports this module.

``sqrn_dbnseq`` mirrors the call signature and return value of the reference's
``SQRNdbnseq`` (/root/reference/src/SQUARNA/SQRNdbnseq.py:973-1286) for the
greedy algorithm with ``bpp 0`` parameter sets, so that the golden vectors
generated from the real reference (tests/golden/make_golden.py) can be compared
field by field.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libsqrn_oracle.so")

GAPS = "-.~"       # seq.py:12
SEPS = ";&"        # seq.py:14

# bracket glyphs, seq.py:108-112
_OPEN = "([{<ABCDEFGHIJKLMNOPQRSTUVWXYZ" + "БГДЁЖЙЛПФЦЧШЩЬЫЪЭЮЯ"
_CLOSE = ")]}>abcdefghijklmnopqrstuvwxyz" + "бгдёжйлпфцчшщьыъэюя"

# reactivity letters, seq.py:17-30
REACT_DICT = {"_": 0.00, "+": 0.50, "#": 1.00, "?": -999}
REACT_DICT.update({"0": 0.05, "1": 0.15, "2": 0.25, "3": 0.35, "4": 0.45,
                   "5": 0.55, "6": 0.65, "7": 0.75, "8": 0.85, "9": 0.95})
REACT_DICT.update({ch: float("%.2f" % (0.04 * k)) for k, ch in enumerate("abcdefghijklmnopqrstuvwxyz")})


def build(force=False):
    """Compile the oracle shared object with oracle/Makefile."""
    src = os.path.join(_HERE, "sqrn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _ParamSet(C.Structure):
    _fields_ = [("nbp", C.c_int), ("keys", C.c_char_p), ("vals", C.POINTER(C.c_double)),
                ("suboptmax", C.c_double), ("suboptmin", C.c_double), ("suboptsteps", C.c_double),
                ("minlen", C.c_double), ("minbpscore", C.c_double), ("minfinscorefactor", C.c_double),
                ("bracketweight", C.c_double), ("distcoef", C.c_double), ("orderpenalty", C.c_double),
                ("loopbonus", C.c_double), ("maxstemnum", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_predict.restype = C.c_void_p
        L.orc_predict.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.POINTER(_ParamSet), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_predict_bpp.restype = C.c_void_p
        L.orc_predict_bpp.argtypes = L.orc_predict.argtypes + [C.c_void_p, C.c_int]
        L.orc_result_count.argtypes = [C.c_void_p]
        L.orc_result_ncalls.argtypes = [C.c_void_p]
        L.orc_result_ncalls.restype = C.c_long
        L.orc_result_nstems.argtypes = [C.c_void_p, C.c_int]
        L.orc_result_get.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 7
        L.orc_result_cons.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_annotate.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                   C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_annotate_bpp.argtypes = L.orc_annotate.argtypes + [C.c_void_p, C.c_int]
        L.orc_optimal.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.POINTER(_ParamSet), C.c_int, C.c_double, C.c_void_p, C.c_int,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]
        L.orc_pair_levels.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_predict_batch_simple.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_ParamSet),
                                               C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_predict_batch_simple.restype = None
        _lib = L
    return _lib


# --------------------------------------------------------------- host helpers
def dbn_to_pairs(dbn):
    """seq.py:172-207: one stack per bracket type, unmatched closers ignored."""
    stacks = {}
    pairs = set()
    for pos, ch in enumerate(dbn):
        k = _OPEN.find(ch)
        if k >= 0:
            stacks.setdefault(k, []).append(pos)
            continue
        k = _CLOSE.find(ch)
        if k >= 0 and stacks.get(k):
            pairs.add((stacks[k].pop(), pos))
    return sorted(pairs)


def unalign(seq, dbn):
    """seq.py:236-255."""
    clean = list(dbn)
    for v, w in dbn_to_pairs(dbn):
        if seq[v] in GAPS or seq[w] in GAPS:
            clean[v] = clean[w] = "."
    keep = [k for k, ch in enumerate(seq) if ch not in GAPS]
    return "".join(seq[k] for k in keep), "".join(clean[k] for k in keep)


def process_reacts(reacts, missing_threshold=-10, middle=0.5, M=1.8, B=1.6):
    """seq.py:32-59 (reverse=False)."""
    neutral = np.exp(-B / M) - 1
    out = []
    for x in reacts or []:
        if x <= missing_threshold or (isinstance(x, float) and math.isnan(x)):
            x = neutral
        else:
            x = min(max(0, x), 1)
        if x <= neutral:
            out.append((middle / neutral) * x)
        else:
            out.append(middle + ((x - neutral) / (1 - neutral)) * (1 - middle))
    return out


def codes_to_dbn(codes):
    out = []
    for c in codes:
        if c == 0:
            out.append(".")
        else:
            lev = abs(int(c)) - 1
            tab = _OPEN if c > 0 else _CLOSE
            out.append(tab[lev] if lev < len(tab) else ".")
    return "".join(out)


def _restraint_arrays(shortrest):
    rclass = np.zeros(max(len(shortrest), 1), dtype=np.uint8)
    for k, ch in enumerate(shortrest):
        if ch in "_+":
            rclass[k] = 1
        elif ch == "/":
            rclass[k] = 2
        elif ch == "\\":
            rclass[k] = 4
    rbps = np.array(dbn_to_pairs(shortrest), dtype=np.int32).reshape(-1, 2)
    return rclass, rbps


class _PS:
    """keeps the ctypes arrays of a list of paramset dicts alive"""

    def __init__(self, paramsets):
        self.keep = []
        self.arr = (_ParamSet * max(len(paramsets), 1))()
        for k, ps in enumerate(paramsets):
            keys = "".join(ps["bpweights"].keys()).encode("latin-1")
            vals = (C.c_double * max(len(ps["bpweights"]), 1))(*ps["bpweights"].values())
            self.keep += [keys, vals]
            self.arr[k] = _ParamSet(len(ps["bpweights"]), keys, vals,
                                    ps["suboptmax"], ps["suboptmin"], ps["suboptsteps"],
                                    ps["minlen"], ps["minbpscore"], ps["minfinscorefactor"],
                                    ps["bracketweight"], ps["distcoef"], ps["orderpenalty"],
                                    ps["loopbonus"], ps["maxstemnum"])
        self.n = len(paramsets)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def predict_short(shortseq, shortreacts, shortrest, paramsets, interchainonly=False, poollim=1000,
                  smat=None, rankby=(0, 2, 1), priority=(), rankbydiff=False, conslim=1,
                  hardrest=False, compensated_sum=False, raw_codes=False, bpp_term=None, bpp_mode=0):
    """Greedy prediction on an ungapped, normalised sequence.  Returns
    (cons_dbn, [ (dbn, scores3, struct_is_int0, [paramset idx], stems[(i,j,len)], bpscores, finscores) ], ncalls).
    raw_codes: the dbns as signed pseudoknot-level codes (+L opening, -L closing, 0 unpaired) instead of strings.
    bpp_term / bpp_mode: the N x N base-pair-probability term of seq.py:341-365 (1: added, 2: multiplied), applied
    for every parameter set of the call."""
    L = lib()
    N = len(shortseq)
    ps = _PS(paramsets)
    reacts = np.ascontiguousarray(shortreacts, dtype=np.float64)
    rclass, rbps = _restraint_arrays(shortrest)
    rk = np.array(rankby, dtype=np.int32)
    pmask = 0
    for p in priority:
        pmask |= 1 << p
    sm = None if smat is None else np.ascontiguousarray(smat, dtype=np.float64)
    bt = None if bpp_term is None else np.ascontiguousarray(bpp_term, dtype=np.float64)
    assert bt is None or bt.shape == (N, N)
    R = L.orc_predict_bpp(shortseq.encode("latin-1"), N, _ptr(reacts), _ptr(rclass), _ptr(rbps), len(rbps),
                          ps.arr, ps.n, int(interchainonly), int(poollim), _ptr(sm), _ptr(rk),
                          pmask, int(rankbydiff), int(conslim), int(hardrest), int(compensated_sum),
                          _ptr(bt), int(bpp_mode) if bt is not None else 0)
    try:
        out = []
        for k in range(L.orc_result_count(R)):
            ns = L.orc_result_nstems(R, k)
            stems = np.zeros((max(ns, 1), 3), dtype=np.int32)
            bpsc = np.zeros(max(ns, 1))
            finsc = np.zeros(max(ns, 1))
            sc = np.zeros(3)
            isint = C.c_int(0)
            mask = C.c_uint64(0)
            dbn = np.zeros(max(N, 1), dtype=np.int32)
            L.orc_result_get(R, k, _ptr(stems), _ptr(bpsc), _ptr(finsc), _ptr(sc), C.byref(isint),
                             C.byref(mask), _ptr(dbn))
            out.append((dbn[:N].astype(np.int8) if raw_codes else codes_to_dbn(dbn[:N]), tuple(float(x) for x in sc), bool(isint.value),
                        [p for p in range(64) if mask.value >> p & 1],
                        [tuple(int(x) for x in row) for row in stems[:ns]],
                        bpsc[:ns].tolist(), finsc[:ns].tolist()))
        cons = np.zeros(max(N, 1), dtype=np.int32)
        L.orc_result_cons(R, _ptr(cons))
        return (cons[:N].astype(np.int8) if raw_codes else codes_to_dbn(cons[:N])), out, L.orc_result_ncalls(R)
    finally:
        L.orc_result_free(R)


def sqrn_dbnseq(seq, reacts=None, restraints=None, dbn=None, paramsets=(), conslim=1, toplim=5,
                hardrest=False, rankbydiff=False, rankby=(0, 2, 1), interchainonly=False,
                poollim=1000, stemmatrix=None, priority=()):
    """Mirror of the reference SQRNdbnseq (greedy sets only): returns
    (consensus_dbn, [(dbn, (total, struct, react), [paramset indices]), ...])."""
    seq = seq.upper().replace("T", "U")                          # seq.py:1004
    if not restraints:
        restraints = "." * len(seq)
    assert len(seq) == len(restraints), "Invalid restraints given"
    if not reacts:
        reacts = [0.5] * len(seq)
    assert len(reacts) == len(seq), "Invalid reactivities given"
    if isinstance(reacts, str):
        reacts = process_reacts([REACT_DICT[ch] for ch in reacts])  # seq.py:1019-1020
    # builtin sum() compensates only exact Python floats (see score_struct in the C file)
    compensated = all(type(x) is float for x in reacts)
    shortseq, shortrest = unalign(seq, restraints)
    keep = [k for k, ch in enumerate(seq) if ch not in GAPS]
    shortreacts = [reacts[k] for k in keep]
    smat = None
    if stemmatrix is not None:
        smat = np.asarray(stemmatrix, dtype=np.float64)[np.ix_(keep, keep)]
    cons, structs, _ = predict_short(shortseq, shortreacts, shortrest, list(paramsets),
                                     interchainonly, poollim, smat, rankby, priority,
                                     rankbydiff, conslim, hardrest, compensated)

    def realign(short):                                           # seq.py:210-233, 1243-1246
        it = iter(short)
        out = []
        for ch in seq:
            if ch in GAPS:
                out.append(".")
            else:
                c = next(it)
                out.append(ch if ch in SEPS else c)
        return "".join(out)

    res = []
    for d, sc, isint, psl, *_ in structs:
        total, struct, react = sc
        res.append((realign(d), (total, 0 if isint else struct, react), psl))
    return realign(cons), res


def annotate(shortseq, paramset, shortreacts=None, shortrest=None, selected=(), interchainonly=False,
             smat=None, bpp_term=None, bpp_mode=0):
    """AnnotateStems seam (seq.py:427-495 after BPMatrix): list of (i, j, len, score)."""
    L = lib()
    N = len(shortseq)
    rclass, rbps = _restraint_arrays(shortrest or "." * N)
    reacts = None if shortreacts is None else np.ascontiguousarray(shortreacts, dtype=np.float64)
    keys = "".join(paramset["bpweights"].keys()).encode("latin-1")
    vals = np.array(list(paramset["bpweights"].values()), dtype=np.float64)
    sel = np.array(list(selected), dtype=np.int32).reshape(-1, 3)
    sm = None if smat is None else np.ascontiguousarray(smat, dtype=np.float64)
    cap = N * N // 2 + 16
    st = np.zeros((cap, 3), dtype=np.int32)
    sc = np.zeros(cap)
    bt = None if bpp_term is None else np.ascontiguousarray(bpp_term, dtype=np.float64)
    assert bt is None or bt.shape == (N, N)
    n = L.orc_annotate_bpp(shortseq.encode("latin-1"), N, _ptr(reacts), _ptr(rclass), _ptr(rbps), len(rbps),
                           len(vals), keys, _ptr(vals), int(interchainonly), float(paramset["minlen"]),
                           float(paramset["minbpscore"]), _ptr(sel), len(sel), _ptr(sm), cap, _ptr(st), _ptr(sc),
                           _ptr(bt), int(bpp_mode) if bt is not None else 0)
    return [(int(st[k, 0]), int(st[k, 1]), int(st[k, 2]), float(sc[k])) for k in range(n)]


def optimal(shortseq, paramset, subopt, shortreacts=None, shortrest=None, selected=(),
            interchainonly=False, smat=None):
    """One OptimalStems call (seq.py:792-833): (scored survivors, chosen)."""
    L = lib()
    N = len(shortseq)
    rclass, rbps = _restraint_arrays(shortrest or "." * N)
    reacts = None if shortreacts is None else np.ascontiguousarray(shortreacts, dtype=np.float64)
    ps = _PS([paramset])
    sel = np.array(list(selected), dtype=np.int32).reshape(-1, 3)
    sm = None if smat is None else np.ascontiguousarray(smat, dtype=np.float64)
    cap = N * N // 2 + 16
    cand = np.zeros((cap, 3), dtype=np.int32)
    cb = np.zeros(cap)
    cf = np.zeros(cap)
    ch = np.zeros((cap, 3), dtype=np.int32)
    chf = np.zeros(cap)
    nc = C.c_int(0)
    n = L.orc_optimal(shortseq.encode("latin-1"), N, _ptr(reacts), _ptr(rclass), _ptr(rbps), len(rbps),
                      ps.arr, int(interchainonly), float(subopt), _ptr(sel), len(sel), _ptr(sm),
                      cap, _ptr(cand), _ptr(cb), _ptr(cf), C.byref(nc), _ptr(ch), _ptr(chf))
    cands = [(int(cand[k, 0]), int(cand[k, 1]), int(cand[k, 2]), float(cb[k]), float(cf[k])) for k in range(nc.value)]
    chosen = [(int(ch[k, 0]), int(ch[k, 1]), int(ch[k, 2]), float(chf[k])) for k in range(n)]
    return cands, chosen


def pair_levels(pairs):
    """PairsToDBN(returnlevels=True), seq.py:104-150: {(v, w): level}."""
    L = lib()
    p = np.array(list(pairs), dtype=np.int32).reshape(-1, 2)
    up = np.zeros((max(len(p), 1), 2), dtype=np.int32)
    lv = np.zeros(max(len(p), 1), dtype=np.int32)
    n = L.orc_pair_levels(_ptr(p), len(p), _ptr(up), _ptr(lv))
    return {(int(up[k, 0]), int(up[k, 1])): int(lv[k]) for k in range(n)}


def predict_batch_simple(seq_bytes, offsets, paramsets, poollim=1, nthreads=1):
    """Config-2 style batch (default reactivities, no restraints): returns
    (dbn level codes int8 in the same CSR layout, scores (B,3), nstems (B,))."""
    L = lib()
    ps = _PS(list(paramsets))
    seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    B = len(offsets) - 1
    dbn = np.zeros(max(len(seq_bytes), 1), dtype=np.int8)
    scores = np.zeros((max(B, 1), 3))
    nst = np.zeros(max(B, 1), dtype=np.int32)
    L.orc_predict_batch_simple(_ptr(seq_bytes), _ptr(offsets), B, ps.arr, ps.n, int(poollim),
                               _ptr(dbn), _ptr(scores), _ptr(nst), int(nthreads))
    return dbn[:len(seq_bytes)], scores[:B], nst[:B]

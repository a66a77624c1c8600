"""Single-sequence prediction: the Python surface of the reference's
SQRNdbnseq.py (same public names, argument order, defaults, return shapes and
printed text), with the greedy hot path running on the GPU through
libsqrn_b200.so.

What stays on the host here is string handling only: normalisation, un/re-
alignment, restraint parsing, reactivity pre-processing, bracket glyphs, metrics
against a known structure and text output.  BPMatrix, AnnotateStems, ScoreStems,
ChooseStems, the structure pool, ScoreStruct and the ranking are behind
``Context.predict_batch`` (csrc/sqrn_abi.cu, csrc/sqrn_device.cuh).

Reference lines cited as seq.py:N are /root/reference/src/SQUARNA/SQRNdbnseq.py.
"""
import math
import sys

import numpy as np

from . import _lib
from ._lib import PackedBatch

GAPS = {'-', '.', '~'}          # seq.py:12
SEPS = {';', '&'}               # seq.py:14

# reactivity letters, seq.py:17-30
ReactDict = {"_": 0.00, "+": 0.50, "#": 1.00, "?": -999}
for _k, _v in zip("0123456789", (0.05, 0.15, 0.25, 0.35, 0.45, 0.55, 0.65, 0.75, 0.85, 0.95)):
    ReactDict[_k] = _v
for _k, _ch in enumerate("abcdefghijklmnopqrstuvwxyz"):
    ReactDict[_ch] = float("%.2f" % (0.04 * _k))

# bracket glyphs per pseudoknot level, seq.py:108-112
_OPEN = "([{<ABCDEFGHIJKLMNOPQRSTUVWXYZ" + "БГДЁЖЙЛПФЦЧШЩЬЫЪЭЮЯ"
_CLOSE = ")]}>abcdefghijklmnopqrstuvwxyz" + "бгдёжйлпфцчшщьыъэюя"
_OPEN_IDX = {c: k for k, c in enumerate(_OPEN)}
_CLOSE_IDX = {c: k for k, c in enumerate(_CLOSE)}

_ctx = {}


def get_context(device=0):
    """Process-wide GPU context (one per device)."""
    if device not in _ctx:
        _ctx[device] = _lib.Context(device)
    return _ctx[device]


# --------------------------------------------------------------------- helpers
def ProcessReacts(reacts, missing_threshold=-10, middle=0.5, reverse=False, M=1.8, B=1.6):
    """Normalise raw reactivities into [0, 1] with the neutral value mapped to
    0.5 (seq.py:32-59).  Values are numpy float64 like the reference's."""
    neutral = np.exp(-B / M) - 1
    if reverse:
        neutral, middle = middle, neutral
    if not reacts:
        return []
    out = []
    for x in reacts:
        if x <= missing_threshold:
            x = neutral
        elif np.isnan(x):
            x = neutral
        else:
            x = min(max(0, x), 1)
        if x <= neutral:
            out.append((middle / neutral) * x)
        else:
            out.append(middle + ((x - neutral) / (1 - neutral)) * (1 - middle))
    return out


def EncodedReactivities(seq, reacts, reactformat):
    """list of floats -> reactivity line (seq.py:82-101)"""
    clipped = [min(max(x, 0), 1) if x == x else 1 for x in reacts]
    if reactformat == 3:
        line = ''.join("_+##"[int(x * 3)] for x in clipped)
    elif reactformat == 10:
        line = ''.join("01234567899"[int(x * 10)] for x in clipped)
    else:
        line = ''.join("abcdefghijklmnopqrstuvwxyz"[int(x * 25 + 0.5)] for x in clipped)
    return ''.join(seq[k] if seq[k] in SEPS else line[k] for k in range(len(seq)))


def DBNToPairs(dbn):
    """dbn string -> sorted list of (i, j); one stack per bracket type, closing
    brackets without a partner are ignored (seq.py:172-207)."""
    stacks = {}
    pairs = set()
    for pos, ch in enumerate(dbn):
        k = _OPEN_IDX.get(ch)
        if k is not None:
            stacks.setdefault(k, []).append(pos)
            continue
        k = _CLOSE_IDX.get(ch)
        if k is not None and stacks.get(k):
            pairs.add((stacks[k].pop(), pos))
    return sorted(pairs)


def _pair_levels(pairs):
    """levels of PairsToDBN (seq.py:119-139) for an arbitrary pair list.
    Returns (unique sorted pairs, insertion order, level per pair index)."""
    ps = sorted(set((min(v, w), max(v, w)) for v, w in pairs))
    n = len(ps)

    def crosses(p, q):
        return (p[0] < q[0] < p[1] < q[1]) or (q[0] < p[0] < q[1] < p[1])

    cc = [sum(1 for b in range(n) if b != a and crosses(ps[a], ps[b])) for a in range(n)]
    order = sorted(range(n), key=lambda a: (cc[a], ps[a][0]))
    groups = []
    for a in order:
        for g in groups:
            if not any(crosses(ps[a], ps[b]) for b in g):
                g.append(a)
                break
        else:
            groups.append([a])
    groups.sort(key=len, reverse=True)
    return ps, groups


def PairsToDBN(newpairs, length=0, returnlevels=False, levellimit=-1):
    """base pairs -> dbn string (seq.py:104-163)"""
    ps, groups = _pair_levels(newpairs)
    if returnlevels:
        return {ps[a]: lev + 1 for lev, g in enumerate(groups) for a in g}
    if levellimit >= 0:
        groups = groups[:levellimit]
    dbn = ['.'] * length
    for lev, g in enumerate(groups):
        op, cl = (_OPEN[lev], _CLOSE[lev]) if lev < len(_OPEN) else ('.', '.')
        for a in g:
            dbn[ps[a][0]] = op
            dbn[ps[a][1]] = cl
    return ''.join(dbn)


def StemsToDBN(stems, seq):
    return PairsToDBN([bp for stem in stems for bp in stem[0]], len(seq))


def ReAlign(shortdbn, longseq, seqmode=False):
    """put the gaps of longseq back into shortdbn (seq.py:210-233)"""
    ngaps = sum(1 for ch in longseq if ch in GAPS)
    assert len(shortdbn) + ngaps == len(longseq), \
        "Cannot ReAlign dbn string - wrong number of gaps:\n{}\n{}".format(longseq, shortdbn)
    it = iter(shortdbn)
    gap = '-' if seqmode else '.'
    return ''.join(gap if ch in GAPS else next(it) for ch in longseq)


def UnAlign(seq, dbn):
    """remove gap columns (and the pairs that touch them) (seq.py:236-255)"""
    clean = list(dbn)
    for v, w in DBNToPairs(dbn):
        if seq[v] in GAPS or seq[w] in GAPS:
            clean[v] = clean[w] = '.'
    keep = [k for k, ch in enumerate(seq) if ch not in GAPS]
    return ''.join(seq[k] for k in keep), ''.join(clean[k] for k in keep)


def ParseRestraints(restraints):
    """restraint string -> (rbps, rxs, rlefts, rrights) (seq.py:370-376)"""
    rbps = DBNToPairs(restraints)
    rxs = {k for k, ch in enumerate(restraints) if ch in ('_', '+')}
    rlefts = {k for k, ch in enumerate(restraints) if ch == '/'}
    rrights = {k for k, ch in enumerate(restraints) if ch == '\\'}
    return rbps, rxs, rlefts, rrights


def PairsToStems(sorted_pairs):
    """group sorted pairs into stacked runs (seq.py:498-517): [[pairs, len], ...]"""
    stems = []
    for k, bp in enumerate(sorted_pairs):
        prev = sorted_pairs[k - 1] if k else None
        if prev is None or not (prev[0] + 1 == bp[0] and prev[1] == bp[1] + 1):
            stems.append([[], 0])
        stems[-1][0].append(bp)
        stems[-1][1] += 1
    return stems


def ScoreStruct(seq, stemset, reacts):
    """The three structure scores of a stem list (seq.py:861-899).  Host copy
    used only for the printed `reference` line; predictions are scored on the
    GPU (team_finalize in csrc/sqrn_device.cuh)."""
    table = {"GU": -0.5, "UG": -0.5, "AU": 1.5, "UA": 1.5, "GC": 4.0, "CG": 4.0}
    thescore = 0
    paired = set()
    for stem in stemset:
        bpsum = 0
        for v, w in stem[0]:
            bpsum += table.get(seq[v] + seq[w], 0.0)
            paired.add(v)
            paired.add(w)
        if bpsum > 0:
            thescore += bpsum ** 1.7
    nonsep = [k for k in range(len(seq)) if seq[k] not in SEPS]
    reactscore = 1 - sum(reacts[k] if k in paired else 1 - reacts[k] for k in nonsep) / len(nonsep)
    return round(thescore * reactscore, 3), round(thescore, 3), round(reactscore, 3)


def ReferenceScores(seq, ref, reacts):
    """scores of the known structure (seq.py:958-970)"""
    if not reacts:
        reacts = [0.5] * len(seq)
    reacts = [reacts[k] for k in range(len(seq)) if seq[k] not in GAPS]
    seq, ref = UnAlign(seq, ref)
    return ScoreStruct(seq, PairsToStems(sorted(DBNToPairs(ref))), reacts)


def _codes_to_dbn(codes):
    """int8 level codes (+L open, -L close) -> glyph string"""
    out = []
    for c in codes.tolist():
        if c == 0:
            out.append('.')
        elif c > 0:
            out.append(_OPEN[c - 1] if c <= len(_OPEN) else '.')
        else:
            out.append(_CLOSE[-c - 1] if -c <= len(_CLOSE) else '.')
    return ''.join(out)


def _metrics(pred, known):
    tp = len(pred & known)
    fp = len(pred - known)
    fn = len(known - pred)
    prc = round(tp / (tp + fp), 3) if (tp + fp) else 1
    rcl = round(tp / (tp + fn), 3) if (tp + fn) else 1
    fsc = round(2 * tp / (2 * tp + fp + fn), 3) if (2 * tp + fp + fn) else 1
    return tp, fp, fn, fsc, prc, rcl


# ------------------------------------------------------------ batch front-end
class _Prepared:
    """one sequence digested the way seq.py:1004-1037 does it"""
    __slots__ = ("seq", "shortseq", "shortrest", "shortreacts", "rbps", "rclass", "keep", "shortdbn",
                 "dbn", "compensated")


def _prepare(seq, reacts, restraints, dbn):
    p = _Prepared()
    seq = seq.upper().replace("T", "U")                               # seq.py:1004
    if not restraints:
        restraints = '.' * len(seq)
    assert len(seq) == len(restraints), "Invalid restraints given"
    if not reacts:
        reacts = [0.5 for _ in range(len(seq))]
    assert len(reacts) == len(seq), "Invalid reactivities given"
    if type(reacts) == str:
        reacts = ProcessReacts([ReactDict[ch] for ch in reacts])      # seq.py:1019-1020 (B = 1.6 default)
    p.seq = seq
    p.shortseq, p.shortrest = UnAlign(seq, restraints)
    p.keep = [k for k, ch in enumerate(seq) if ch not in GAPS]
    p.shortreacts = [reacts[k] for k in p.keep]
    # builtin sum() in ScoreStruct compensates exact Python floats only (CPython >= 3.12)
    p.compensated = all(type(x) is float for x in p.shortreacts)
    p.dbn = dbn
    p.shortdbn = None
    if dbn:
        assert len(seq) == len(dbn)
        p.shortdbn = UnAlign(seq, dbn)[1]
    rbps, rxs, rlefts, rrights = ParseRestraints(p.shortrest)
    p.rbps = rbps
    rc = np.zeros(max(len(p.shortseq), 1), dtype=np.uint8)
    for k in rxs:
        rc[k] |= 1
    for k in rlefts:
        rc[k] |= 2
    for k in rrights:
        rc[k] |= 4
    p.rclass = rc[:len(p.shortseq)]
    return p


def _encode_symbols(shortseq):
    """str -> bytes for the C ABI; non-latin-1 symbols can never pair and become '?'"""
    return shortseq.encode("latin-1", "replace")


def predict_many(entries, paramsets, conslim=1, toplim=5, hardrest=False, rankbydiff=False,
                 rankby=(0, 2, 1), interchainonly=False, stemmatrix=None, poollim=1000,
                 priority=frozenset(), algos=frozenset(), device=0):
    """Batched SQRNdbnseq: entries = [(seq, reacts, restraints, dbn)], one GPU
    call for all of them.  Returns the reference's 4-tuple per entry."""
    assert set(rankby) == {0, 1, 2} and len(rankby) == 3, "Invalid ranking indices"
    # which parameter sets run the greedy algorithm (seq.py:1046-1102)
    gsets = []
    for psi, ps in enumerate(paramsets):
        use = set(algos) if algos else set(ps["algorithms"])
        if ps.get("bpp", 0):
            raise NotImplementedError("parameter set #{} has bpp != 0: ViennaRNA base-pair probabilities are "
                                      "outside the GPU hot path (use the *nobpp configs)".format(psi))
        if use - {"G"}:
            raise NotImplementedError("algorithms {} are outside the GPU hot path (greedy 'G' only)"
                                      .format(sorted(use - {"G"})))
        if "G" in use:
            gsets.append(psi)
    preps = [_prepare(*e) for e in entries]
    results = [None] * len(entries)
    if not gsets:
        todo = []
    else:
        todo = list(range(len(entries)))
    # sequences sharing the same reactivity-sum mode go into the same batch
    for comp in (False, True):
        idx = [k for k in todo if preps[k].compensated == comp]
        if not idx:
            continue
        any_react = any(any(x != 0.5 for x in preps[k].shortreacts) for k in idx)
        codes = values = None
        if any_react:
            table = {}
            codes = []
            for k in idx:
                arr = np.empty(len(preps[k].shortreacts), dtype=np.uint16)
                for q, x in enumerate(preps[k].shortreacts):
                    x = float(x)
                    c = table.get(x)
                    if c is None:
                        c = table[x] = len(table)
                    arr[q] = c
                codes.append(arr)
            if len(table) > 65535:
                raise NotImplementedError("more than 65535 distinct reactivity values in one batch")
            values = np.array(list(table.keys()), dtype=np.float64)
        any_restr = any(p.rbps or p.rclass.any() for p in (preps[k] for k in idx))
        smat = cols = None
        if stemmatrix is not None:
            smat = np.asarray(stemmatrix, dtype=np.float64)
            cols = [np.array(preps[k].keep, dtype=np.int32) for k in idx]
        pmask = 0
        for p in priority:
            if p in gsets:
                pmask |= 1 << gsets.index(p)
        batch = PackedBatch([_encode_symbols(preps[k].shortseq) for k in idx],
                            react_codes=codes, react_values=values, react_comp=comp,
                            restr_class=[preps[k].rclass for k in idx] if any_restr else None,
                            rbps=[np.array(preps[k].rbps, dtype=np.int32).reshape(-1, 2) for k in idx] if any_restr else None,
                            smat=smat, cols=cols, interchainonly=interchainonly, hardrest=hardrest,
                            rankbydiff=rankbydiff, poollim=poollim, conslim=conslim, max_structs=0,
                            rankby=rankby, priority_mask=pmask)
        out = get_context(device).predict_batch([paramsets[g] for g in gsets], batch)
        for k, (cons, structs, _ntot) in zip(idx, out):
            results[k] = (cons, structs)

    final = []
    for k, p in enumerate(preps):
        seq = p.seq

        def expand(short):                      # ReAlign + separators, seq.py:1239-1246
            long_ = ReAlign(short, seq)
            return ''.join(seq[q] if seq[q] in SEPS else long_[q] for q in range(len(seq)))

        if results[k] is None:
            cons_codes, structs = np.zeros(len(p.shortseq), np.int8), []
        else:
            cons_codes, structs = results[k]
        cons = expand(_codes_to_dbn(cons_codes))
        preds = []
        bpsets = []
        for codes, sc, isint, mask, stems in structs:
            total, struct, react = sc
            inds = [gsets[b] for b in range(len(gsets)) if mask >> b & 1]
            preds.append((expand(_codes_to_dbn(codes)), (total, 0 if isint else struct, react), inds))
            bpsets.append(codes)
        if p.dbn:                                # seq.py:1249-1285
            known = set(DBNToPairs(p.shortdbn))
            consresult = list(_metrics(set(DBNToPairs(_codes_to_dbn(cons_codes))), known))
            best, result = -1, []
            for rank, codes in enumerate(bpsets):
                tp, fp, fn, fsc, prc, rcl = _metrics(set(DBNToPairs(_codes_to_dbn(codes))), known)
                if fsc > best:
                    best = fsc
                    result = [tp, fp, fn, fsc, prc, rcl, rank + 1]
                if rank + 1 >= toplim:
                    break
            final.append((cons, preds, consresult, result))
        else:
            final.append((cons, preds, [np.nan] * 6, [np.nan] * 7))
    return final


def SQRNdbnseq(seq, reacts=None, restraints=None, dbn=None,
               paramsets=[], conslim=1, toplim=5,
               hardrest=False, rankbydiff=False,
               rankby=(0, 2, 1), interchainonly=False,
               threads=1, mp=True, stemmatrix=None, poollim=1000,
               entropy=False, algos=set(), levellimit=None,
               priority=set(),
               M=1.8, B=-0.6):
    """Predict alternative secondary structures of one sequence; same signature
    and return value as the reference (seq.py:973-1286):
    (consensus_dbn, [(dbn, (total, struct, react), [paramset indices]), ...],
     consensus_metrics[6], topN_metrics[7]).

    threads / mp only choose a multiprocessing layout in the reference and never
    change results; here the work is one batched GPU call.  `entropy` and the
    non-greedy algorithms are outside the GPU hot path (NotImplementedError)."""
    if entropy:
        raise NotImplementedError("entropy mode is outside the GPU hot path")
    return predict_many([(seq, reacts, restraints, dbn)], paramsets, conslim, toplim, hardrest,
                        rankbydiff, rankby, interchainonly, stemmatrix, poollim,
                        frozenset(priority), frozenset(algos))[0]


def _print_entry(name, sequence, reactivities, restraints, reference, reactformat, sink, rfam=None):
    """header block of RunSQRNdbnseq (seq.py:1301-1345)"""
    print(name, file=sink)
    print(sequence, file=sink)
    if reactivities:
        print(EncodedReactivities(sequence, reactivities, reactformat), "reactivities", sep='\t', file=sink)
    if restraints:
        print(''.join(sequence[k] if sequence[k] in SEPS else restraints[k] for k in range(len(sequence))),
              "restraints" + ("(" + rfam + ")" if rfam else ""), sep='\t', file=sink)
    if reference:
        print(''.join(sequence[k] if sequence[k] in SEPS else reference[k] for k in range(len(sequence))),
              "reference", *ReferenceScores(sequence, reference, reactivities), sep='\t', file=sink)
    print('_' * len(sequence), file=sink)


def _print_prediction(prediction, sequence, restraints, reference, paramsetnames, conslim, outplim, sink,
                      rfam=None):
    """result block of RunSQRNdbnseq (seq.py:1357-1406)"""
    consensus, predicted_structures, consensus_metrics, topN_metrics = prediction
    g4 = bool(rfam and restraints and '+' in restraints)
    if g4:
        consensus = ''.join('+' if restraints[k] == '+' else ch for k, ch in enumerate(consensus))
    if reference:
        print(consensus, "top-{}_consensus".format(conslim),
              "TP={},FP={},FN={},FS={},PR={},RC={}".format(*consensus_metrics), sep='\t', file=sink)
    else:
        print(consensus, "top-{}_consensus".format(conslim), sep='\t', file=sink)
    print('=' * len(sequence), file=sink)
    for k, (struct, scores, inds) in enumerate(predicted_structures[:outplim]):
        if g4:
            struct = ''.join('+' if restraints[q] == '+' else ch for q, ch in enumerate(struct))
        total, structscore, reactscore = scores
        fields = [struct, "#{}".format(k + 1), total, structscore, reactscore,
                  ','.join(paramsetnames[q] for q in inds)]
        if reference and k + 1 == topN_metrics[-1]:
            fields.append("TP={},FP={},FN={},FS={},PR={},RC={},RK={}".format(*topN_metrics))
        print(*fields, sep='\t', file=sink)
    return consensus, predicted_structures, consensus_metrics, topN_metrics


def _resolve_priority(priority, paramsetnames, rfam):
    if rfam and priority == {'bppN', 'bppH1', 'bppH2'}:          # seq.py:1304-1305
        priority = None
    if priority:
        return {k for k in range(len(paramsetnames)) if paramsetnames[k] in priority}
    return set()


def RunSQRNdbnseq(name, sequence, reactivities, restraints,
                  reference, paramsetnames,
                  paramsets, threads, rankbydiff, rankby,
                  hardrest, interchainonly, toplim, outplim,
                  conslim, reactformat, evalonly, poollim=1000,
                  mp=True, sink=sys.stdout, stemmatrix=None,
                  entropy=False, algos={'G', }, levellimit=None,
                  priority=None, rfam=None, M=1.8, B=-0.6):
    """Print the reference's text block for one entry and return the prediction
    4-tuple (seq.py:1289-1408)."""
    if entropy:
        raise NotImplementedError("entropy mode is outside the GPU hot path")
    priority = _resolve_priority(priority, paramsetnames, rfam)
    _print_entry(name, sequence, reactivities, restraints, reference, reactformat, sink, rfam)
    if evalonly:
        return None, None, None, None
    prediction = SQRNdbnseq(sequence, reactivities, restraints, reference,
                            paramsets, conslim, toplim, hardrest,
                            rankbydiff, rankby, interchainonly, threads, mp, stemmatrix,
                            poollim, algos=algos, levellimit=levellimit, priority=priority,
                            M=M, B=B)
    return _print_prediction(prediction, sequence, restraints, reference, paramsetnames, conslim, outplim,
                             sink, rfam)


def RunSQRNdbnseqBatch(entries, paramsetnames, paramsets, rankbydiff, rankby, hardrest, interchainonly,
                       toplim, outplim, conslim, reactformat, evalonly, poollim=1000, sink=sys.stdout,
                       stemmatrix=None, algos={'G', }, priority=None, rfam=None):
    """RunSQRNdbnseq for many entries [(name, seq, reacts, restraints, reference)]
    with ONE batched GPU call; text is written in input order (what the
    reference's ordered imap gives, SQUARNA.py:929-935)."""
    priority = _resolve_priority(priority, paramsetnames, rfam)
    preds = [None] * len(entries)
    if not evalonly:
        preds = predict_many([(e[1], e[2], e[3], e[4]) for e in entries], paramsets, conslim, toplim,
                             hardrest, rankbydiff, rankby, interchainonly, stemmatrix, poollim,
                             frozenset(priority), frozenset(algos))
    out = []
    for (name, seq, reacts, rests, ref), pred in zip(entries, preds):
        _print_entry(name, seq, reacts, rests, ref, reactformat, sink, rfam)
        if evalonly:
            out.append((None, None, None, None))
        else:
            out.append(_print_prediction(pred, seq, rests, ref, paramsetnames, conslim, outplim, sink, rfam))
    return out

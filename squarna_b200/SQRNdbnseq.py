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


# code point of every int8 level code (+L opening bracket of level L, -L closing one), indexed by
# the code reinterpreted as uint8; levels beyond the 49 bracket pairs print as '.' (seq.py:142-143)
_GLYPH32 = np.full(256, ord('.'), dtype=np.uint32)
for _lev in range(1, 128):
    if _lev <= len(_OPEN):
        _GLYPH32[_lev] = ord(_OPEN[_lev - 1])
        _GLYPH32[256 - _lev] = ord(_CLOSE[_lev - 1])


def _codes_to_dbn(codes):
    """int8 level codes (+L open, -L close) -> glyph string"""
    return _GLYPH32[np.asarray(codes, dtype=np.int8).view(np.uint8)].tobytes().decode("utf-32-le")


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
    __slots__ = ("seq", "shortseq", "shortrest", "_sr", "rbps", "rclass", "_keep", "shortdbn",
                 "dbn", "compensated")

    @property
    def shortreacts(self):
        """processed reactivities of the ungapped sequence (all 0.5 when none were given)"""
        return [0.5] * len(self.shortseq) if self._sr is None else self._sr

    @property
    def keep(self):
        """ungapped position -> position in the input sequence"""
        return np.arange(len(self.seq)) if self._keep is None else self._keep


_REACT_LUT = {}


def _react_lut(M=1.8, B=1.6):
    """processed reactivity of every reactivity letter (seq.py:1019-1020 -> ProcessReacts defaults)"""
    key = (M, B)
    if key not in _REACT_LUT:
        chars = sorted(ReactDict)
        vals = ProcessReacts([ReactDict[ch] for ch in chars], M=M, B=B)
        lut = np.full(256, np.nan)
        for ch, v in zip(chars, vals):
            lut[ord(ch)] = v
        _REACT_LUT[key] = lut
    return _REACT_LUT[key]


_GAP_BYTES = np.zeros(256, dtype=bool)
for _ch in GAPS:
    _GAP_BYTES[ord(_ch)] = True
_RC_BYTES = np.zeros(256, dtype=np.uint8)
_RC_BYTES[ord('_')] = _RC_BYTES[ord('+')] = 1
_RC_BYTES[ord('/')] = 2
_RC_BYTES[ord('\\')] = 4
_PLAIN_RESTR = np.zeros(256, dtype=bool)           # restraint symbols that are not brackets
for _ch in "._+/\\-~":
    _PLAIN_RESTR[ord(_ch)] = True


def _prepare(seq, reacts, restraints, dbn):
    """seq.py:1004-1037 for one entry.  The common shapes (no gaps, no restraints, encoded or absent
    reactivities) take vectorised paths; everything else goes through the reference's own steps."""
    p = _Prepared()
    seq = seq.upper().replace("T", "U")                               # seq.py:1004
    n = len(seq)
    if restraints:
        assert n == len(restraints), "Invalid restraints given"
    if reacts:
        assert len(reacts) == n, "Invalid reactivities given"
    p.seq = seq
    raw = np.frombuffer(seq.encode("latin-1", "replace"), dtype=np.uint8)
    gapmask = _GAP_BYTES[raw]
    has_gaps = bool(gapmask.any())
    p._keep = np.flatnonzero(~gapmask) if has_gaps else None         # None: identity
    # --- restraints -----------------------------------------------------------
    if not restraints:
        p.shortseq = ''.join(seq[k] for k in p._keep) if has_gaps else seq
        p.shortrest = '.' * len(p.shortseq)
        p.rbps = []
        p.rclass = np.zeros(len(p.shortseq), dtype=np.uint8)
    else:
        rraw = np.frombuffer(restraints.encode("latin-1", "replace"), dtype=np.uint8)
        if not has_gaps and bool(_PLAIN_RESTR[rraw].all()):
            p.shortseq, p.shortrest, p.rbps = seq, restraints, []    # no brackets: nothing for DBNToPairs
            p.rclass = _RC_BYTES[rraw]
        else:
            p.shortseq, p.shortrest = UnAlign(seq, restraints)
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
    # --- reactivities ---------------------------------------------------------
    if not reacts:
        p._sr = None                                                  # all 0.5: "default reacts"
        p.compensated = True
    elif type(reacts) == str:
        vals = _react_lut()[np.frombuffer(reacts.encode("latin-1", "replace"), dtype=np.uint8)]
        if np.isnan(vals).any():
            raise KeyError(next(ch for ch in reacts if ch not in ReactDict))
        p._sr = vals[p._keep] if has_gaps else vals                   # numpy floats, like ProcessReacts' output
        p.compensated = False
    else:
        keep = p._keep if has_gaps else range(n)
        p._sr = [reacts[k] for k in keep]
        # builtin sum() in ScoreStruct compensates exact Python floats only (CPython >= 3.12)
        p.compensated = all(type(x) is float for x in p._sr)
    p.dbn = dbn
    p.shortdbn = None
    if dbn:
        assert len(seq) == len(dbn)
        p.shortdbn = UnAlign(seq, dbn)[1]
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
        # distinct processed reactivities of the batch -> codes + value table (host pow() table in the library)
        codes = values = None
        arrs = [None if preps[k]._sr is None else np.asarray(preps[k]._sr, dtype=np.float64) for k in idx]
        if any(a is not None and bool((a != 0.5).any()) for a in arrs):
            lens_ = [len(preps[k].shortseq) for k in idx]
            flat = np.concatenate([np.full(n_, 0.5) if a is None else a for a, n_ in zip(arrs, lens_)]) if idx else np.zeros(0)
            # bit patterns, not values: -0.0 / NaN payloads must stay distinct table entries
            values_bits, inverse = np.unique(flat.view(np.uint64), return_inverse=True)
            if len(values_bits) > 65535:
                raise NotImplementedError("more than 65535 distinct reactivity values in one batch")
            values = values_bits.view(np.float64)
            inverse = inverse.astype(np.uint16)
            codes, o = [], 0
            for n_ in lens_:
                codes.append(inverse[o:o + n_])
                o += n_
        any_restr = any(p.rbps or p.rclass.any() for p in (preps[k] for k in idx))
        smat = cols = None
        if stemmatrix is not None:
            smat = np.asarray(stemmatrix, dtype=np.float64)
            cols = [np.asarray(preps[k].keep, dtype=np.int32) for k in idx]
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

        raw = np.frombuffer(seq.encode("utf-32-le"), dtype=np.uint32)
        seppos = np.flatnonzero((raw == ord(';')) | (raw == ord('&')))
        keep = p._keep

        def expand(codes, raw=raw, seppos=seppos, keep=keep):     # glyphs + ReAlign + separators, seq.py:1239-1246
            g = _GLYPH32[np.asarray(codes, dtype=np.int8).view(np.uint8)]
            if keep is not None:
                long_ = np.full(len(raw), ord('.'), dtype=np.uint32)
                long_[keep] = g
                g = long_
            if len(seppos):
                g = g.copy() if keep is None else g
                g[seppos] = raw[seppos]
            return g.tobytes().decode("utf-32-le")

        if results[k] is None:
            cons_codes, structs = np.zeros(len(p.shortseq), np.int8), []
        else:
            cons_codes, structs = results[k]
        cons = expand(cons_codes)
        preds = []
        bpsets = []
        for codes, sc, isint, mask, stems in structs:
            total, struct, react = sc
            inds = [gsets[b] for b in range(len(gsets)) if mask >> b & 1]
            preds.append((expand(codes), (total, 0 if isint else struct, react), inds))
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

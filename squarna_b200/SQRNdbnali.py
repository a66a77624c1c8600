"""Alignment mode: the surface of the reference's SQRNdbnali.py.

GPU work: step 1 enumerates the stems of every sequence (YieldStems =
BPMatrix + AnnotateStems) in one batched call of sqrn_yield_stems_batch; step 2
runs the stem-matrix-weighted single-sequence predictions of all sequences in
one batched sqrn_predict_batch call.  What stays on the host, in sequence order
because float64 accumulation order is observable: summing stem
scores into the L x L matrix, MatrixToDBNs, Consensus and the text output.

Lines cited as ali.py:N are /root/reference/src/SQUARNA/SQRNdbnali.py.
"""
import io
import os
import sys

import numpy as np

from . import SQRNdbnseq as _seq
from ._lib import PackedBatch
from .SQRNdbnseq import (DBNToPairs, EncodedReactivities, GAPS, PairsToDBN, SEPS, UnAlign)


def ReAlignDict(shortseq, longseq):
    """ungapped index -> aligned column (ali.py:20-37)"""
    cols = [k for k, ch in enumerate(longseq) if ch not in GAPS]
    return {k: cols[k] for k in range(len(shortseq))}


def _yield_many(entries, bpweights, interchainonly, minlen, minbpscore, device=None, matrix=None):
    """entries: [(seq, reacts or None, restraints or None)] (aligned).  One GPU call per GPU: with several visible
    devices (device=None) the rows are dealt to one host thread per GPU (the reference's Pool.imap over sequences,
    ali.py:222-233) and come back in input order.
    Returns per entry (cols int32[n_ungapped], stems int32[k,3] ungapped, scores float64[k])."""
    if device is None:
        devs = _seq._resolve_devices(None)
        if len(devs) > 1 and len(entries) >= 8 * len(devs) and matrix is None:
            return _seq.run_sharded(lambda sub, dev: _yield_many(sub, bpweights, interchainonly, minlen, minbpscore, dev),
                                    entries, [len(e[0]) for e in entries], devs, exponent=2.0)
        device = devs[0]
    gap_bytes = np.frombuffer("".join(sorted(GAPS)).encode("latin-1"), dtype=np.uint8)
    ps = dict(algorithms={"G"}, bpp=0.0, bpweights=bpweights, suboptmax=1.0, suboptmin=1.0, suboptsteps=1.0,
              minlen=minlen, minbpscore=minbpscore, minfinscorefactor=1.0, distcoef=0.0, bracketweight=-2.0,
              orderpenalty=0.0, loopbonus=0.0, maxstemnum=1e6)
    if matrix is not None and len(entries) >= 16 and all(len(e[0]) == matrix[0] for e in entries):
        batch = _rows_batch(entries, matrix[0], gap_bytes, interchainonly)
        if batch is not None:
            return _seq.get_context(device).stem_matrix(ps, batch, matrix[1])
    preps = []
    for seq, reacts, rests in entries:
        seq = seq.upper().replace("T", "U")                           # ali.py:65
        if not rests:
            rests = '.' * len(seq)
        assert len(seq) == len(rests)
        raw = np.frombuffer(seq.encode("latin-1", "replace"), dtype=np.uint8)
        if len(raw) != len(seq):                                       # (cannot happen: latin-1 is one byte per symbol)
            raise ValueError("sequence encoding")
        keep = np.flatnonzero(~np.isin(raw, gap_bytes)).astype(np.int32)
        if rests.count('.') == len(rests):
            # no restraints (the usual row of an alignment): nothing to parse, UnAlign is the column selection
            shortseq, rbps = raw[keep].tobytes(), []
            rc = np.zeros(len(keep), dtype=np.uint8)
        else:
            shortseq, shortrest = UnAlign(seq, rests)
            rbps, rxs, rlefts, rrights = _seq.ParseRestraints(shortrest)
            rc = np.zeros(max(len(shortseq), 1), dtype=np.uint8)
            for k in rxs:
                rc[k] |= 1
            for k in rlefts:
                rc[k] |= 2
            for k in rrights:
                rc[k] |= 4
            rc = rc[:len(shortseq)]
            shortseq = _seq._encode_symbols(shortseq)
        shortreacts = np.asarray(reacts, dtype=np.float64)[keep] if reacts else None    # ali.py:77-82
        preps.append((shortseq, shortreacts, rbps, rc, keep))
    any_react = any(p[1] is not None and bool((p[1] != 0.5).any()) for p in preps)
    codes = values = None
    if any_react:
        # distinct reactivity values of the batch -> value table + one code per position
        flat = np.concatenate([p[1] if p[1] is not None else np.full(len(p[0]), 0.5) for p in preps])
        values, inv = np.unique(flat, return_inverse=True)
        if len(values) > 65535:
            raise ValueError("more than 65535 distinct reactivity values in one alignment")
        inv = inv.astype(np.uint16)
        codes, at = [], 0
        for p in preps:
            codes.append(inv[at:at + len(p[0])])
            at += len(p[0])
        values = np.ascontiguousarray(values, dtype=np.float64)
    any_restr = any(p[2] or p[3].any() for p in preps)
    batch = PackedBatch([p[0] for p in preps], react_codes=codes, react_values=values,
                        restr_class=[p[3] for p in preps] if any_restr else None,
                        rbps=[np.array(p[2], dtype=np.int32).reshape(-1, 2) for p in preps] if any_restr else None,
                        interchainonly=interchainonly,
                        cols=[p[4] for p in preps] if matrix is not None else None,
                        ali_len=matrix[0] if matrix is not None else 0)
    if matrix is not None:
        # the whole of step 1 on the device: stems, the sequence-ordered sum into the L x L matrix, the cells MatrixToDBNs
        # walks (sqrn_stem_matrix_batch); matrix = (alignment length, score threshold)
        return _seq.get_context(device).stem_matrix(ps, batch, matrix[1])
    out = _seq.get_context(device).yield_stems(ps, batch)
    return [(p[4], st, sc) for p, (st, sc) in zip(preps, out)]


def _rows_batch(entries, L, gap_bytes, interchainonly):
    """The rows of an alignment as one CSR batch without a Python loop over the rows: UnAlign (ali.py:70-82) is a column
    selection per row, a restraint line shared by the rows is parsed once in aligned coordinates (a pair survives in a row
    when both of its columns hold a symbol there: seq.py:236-255), reactivity lists shared by the rows are indexed by column.
    None: the rows differ in their restraint lines or reactivity lists in a way this path does not cover."""
    n = len(entries)
    rest_lines = {e[2] if e[2] else None for e in entries}
    if len(rest_lines) > 1:
        return None
    rests = next(iter(rest_lines))
    react_ids = {id(e[1]) if e[1] else 0 for e in entries}
    if len(react_ids) > 1:
        first = next(e[1] for e in entries if e[1])
        if not all(e[1] == first for e in entries if e[1]) or any(not e[1] for e in entries):
            return None
    reacts = next((e[1] for e in entries if e[1]), None)
    text = "".join(e[0].upper().replace("T", "U") for e in entries)
    raw = np.frombuffer(text.encode("latin-1", "replace"), dtype=np.uint8)
    if raw.size != n * L:
        return None
    raw = raw.reshape(n, L)
    mask = ~np.isin(raw, gap_bytes)
    lens = mask.sum(axis=1)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    rows_idx, cols_flat = np.nonzero(mask)                      # row-major: the CSR order
    symbols = raw[mask]
    kw = {}
    if rests is not None and rests.count('.') != len(rests):
        if len(rests) != L:
            return None
        idx = np.cumsum(mask, axis=1, dtype=np.int32) - 1       # ungapped index of every column, per row
        pairs = np.array(DBNToPairs(rests), dtype=np.int64).reshape(-1, 2)
        cls_line = np.zeros(L, np.uint8)
        for k, ch in enumerate(rests):
            if ch in '_+':
                cls_line[k] = 1
            elif ch == '/':
                cls_line[k] = 2
            elif ch == '\\':
                cls_line[k] = 4
        if len(pairs):
            alive = mask[:, pairs[:, 0]] & mask[:, pairs[:, 1]]  # n x P
            r_, p_ = np.nonzero(alive)
            rbps = np.stack([idx[r_, pairs[p_, 0]], idx[r_, pairs[p_, 1]]], axis=1).astype(np.int32)
            rb_off = np.zeros(n + 1, np.int64)
            np.cumsum(alive.sum(axis=1), out=rb_off[1:])
        else:
            rbps, rb_off = np.zeros((0, 2), np.int32), np.zeros(n + 1, np.int64)
        kw["restr_class"] = cls_line[cols_flat]
        kw["rbps"] = (rb_off, rbps)
    if reacts is not None:
        arr = np.asarray(reacts, dtype=np.float64)
        if arr.shape != (L,):
            return None
        if bool((arr != 0.5).any()):
            values, inv = np.unique(arr, return_inverse=True)
            kw["react_codes"] = inv.astype(np.uint16)[cols_flat]
            kw["react_values"] = np.ascontiguousarray(values, np.float64)
    return PackedBatch((symbols, offsets), interchainonly=interchainonly, cols=cols_flat.astype(np.int32), ali_len=L,
                       flat=True, **kw)


def YieldStems(seq, reactivities=None, restraints=None,
               bpweights={}, interchainonly=False,
               minlen=2, minbpscore=0, M=1.8, B=-0.6):
    """stems of one (possibly aligned) sequence in aligned coordinates:
    [[[(v, w), ...], score], ...] in the reference's order (ali.py:60-108)"""
    cols, stems, scores = _yield_many([(seq, reactivities, restraints)], bpweights, interchainonly,
                                      minlen, minbpscore)[0]
    out = []
    for (i, j, ln), sc in zip(stems.tolist(), scores.tolist()):
        out.append([[(int(cols[i + k]), int(cols[j - k])) for k in range(ln)], sc])
    return out


def MatrixToDBNs(mat, score, depth, verbose=False, sink=sys.stdout, cells=None):
    """greedy assembly of structures from the stem-score matrix (ali.py:121-192):
    cells in stable descending order, stop below score*depth, accept w - v >= 4,
    first-fit into position-disjoint structures.  cells: the flat indices of the cells that pass both tests, already
    in that order (what sqrn_stem_matrix_batch returns); None: sorted here."""
    N = mat.shape[0]
    thr = score * depth
    flat = mat.ravel()
    if cells is None:
        order = np.argsort(-flat, kind="stable")      # ties keep ascending flat index, like sorted(reverse=True)
    else:
        order = np.asarray(cells)
    res = [[[], set()]]
    if verbose:
        print(">Conserved base pairs (one by one)", file=sink)
    for idx in order.tolist():
        val = flat[idx]
        if val < thr:
            break
        v, w = divmod(idx, N)
        if not w - v >= 4:
            continue
        for struct in res:
            if v not in struct[1] and w not in struct[1]:
                struct[0].append((v, w))
                struct[1].update((v, w))
                break
        else:
            res.append([[(v, w)], {v, w}])
        if verbose:
            print(PairsToDBN([(v, w)], N), round(val, 3), sep='\t', file=sink)
    dbns = [PairsToDBN(struct[0], N) for struct in res]
    if verbose:
        print(">Conserved base pairs (assembled)", file=sink)
        for dbn in dbns:
            print(dbn, file=sink)
    return dbns


def Metrics(ref, pred):
    """TP, FP, FN, FS, PR, RC of two dbn strings (ali.py:195-208)"""
    if not ref:
        return [np.nan] * 6
    return list(_seq._metrics(set(DBNToPairs(pred)), set(DBNToPairs(ref))))


def SQRNdbnali(objs, defrests=None, defreacts=None, defref=None,
               bpweights={}, interchainonly=False,
               minlen=2, minbpscore=0,
               threads=1, verbose=False,
               sink=sys.stdout, M=1.8, B=-0.6):
    """step 1 of the alignment mode: (predicted dbn, stem matrix) (ali.py:211-242)"""
    L = len(objs[0][1])
    entries = [(obj[1], obj[2], defrests if defrests else obj[3]) for obj in objs]
    if len(_seq._resolve_devices(None)) == 1 and not os.environ.get("SQRN_HOST_STEMMATRIX"):
        # one GPU: stems, sequence-ordered accumulation and the ranking of the conserved cells all stay on the device
        stemmatrix, cells = _yield_many(entries, bpweights, interchainonly, minlen, minbpscore,
                                        matrix=(L, minbpscore * len(objs)))
        pred_dbns = MatrixToDBNs(stemmatrix, minbpscore, len(objs), verbose, sink=sink, cells=cells)
        return pred_dbns[0], stemmatrix
    # several GPUs: every GPU enumerates the stems of its share of the rows; the only exchange of the alignment mode --
    # the sum into the matrix -- is done here on the host, in sequence order (float64 addition order is observable)
    stemmatrix = _accumulate_host(_yield_many(entries, bpweights, interchainonly, minlen, minbpscore), L)
    pred_dbns = MatrixToDBNs(stemmatrix, minbpscore, len(objs), verbose, sink=sink)
    return pred_dbns[0], stemmatrix


def _accumulate_host(per_seq, L):
    """the stem-score matrix from per-sequence stems (cols, stems, scores): sequence order, stem order, outer->inner
    pairs -- the reference's accumulation order (ali.py:233-237).  Within one sequence every cell is touched by at most
    one stem, so a fancy-indexed add per sequence performs exactly the same float64 additions per cell."""
    stemmatrix = np.zeros((L, L))
    for cols, stems, scores in per_seq:
        if not len(stems):
            continue
        lens = stems[:, 2]
        rep = np.repeat(np.arange(len(stems)), lens)
        k = np.arange(lens.sum()) - np.repeat(np.cumsum(lens) - lens, lens)
        v = cols[stems[rep, 0] + k]
        w = cols[stems[rep, 1] - k]
        s = scores[rep]
        stemmatrix[v, w] += s
        stemmatrix[w, v] += s
    return stemmatrix


def Consensus(structs, freqlimit=0.0, verbose=False, sink=sys.stdout):
    """most frequent non-conflicting pairs of a list of dbns (ali.py:271-304)"""
    counts = {}
    limit = freqlimit * len(structs)
    for struct in structs:
        for bp in DBNToPairs(struct):
            counts[bp] = counts.get(bp, 0) + 1
    chosen, seen = [], set()
    if verbose:
        print(">Step 2, Populated base pairs", file=sink)
    for bp in sorted(counts, key=lambda x: counts[x], reverse=True):      # stable: first-seen order on ties
        if verbose:
            print(PairsToDBN([bp], len(structs[0])), counts[bp], file=sink)
        if counts[bp] >= limit and bp[0] not in seen and bp[1] not in seen:
            seen.update(bp)
            chosen.append(bp)
    return PairsToDBN(list(set(chosen)), len(structs[0]))


def ReactScore(reacts, seq, dbn):
    """1 - mean reactivity error of a structure (ali.py:307-329)"""
    if not reacts:
        return 0.5
    paired = {p for bp in DBNToPairs(dbn) for p in bp}
    nonsep = [k for k in range(len(seq)) if seq[k] not in SEPS]
    return 1 - sum(reacts[k] if k in paired else 1 - reacts[k] for k in nonsep) / len(nonsep)


def RunSQRNdbnali(objs, defreacts, defrests, defref,
                  levellimit, freqlimit, verbose, step3,
                  paramsetnames, paramsets, threads, rankbydiff, rankby,
                  hardrest, interchainonly, toplim, outplim,
                  conslim, reactformat, poollim, entropy=False,
                  algos={'G', }, sink=sys.stdout, M=1.8, B=-0.6):
    """the three-step alignment-based prediction and its printed report (ali.py:332-458)"""
    N = len(objs[0][1])
    first = paramsets[0]
    bpweights, minlen, minbpscore = first['bpweights'], first['minlen'], first['minbpscore']

    if verbose:
        print(">Step 1, Iteration 1", file=sink)
    pred_dbn, smat = SQRNdbnali(objs, defrests, defreacts, defref, bpweights, interchainonly, minlen, minbpscore,
                                threads, verbose, sink=sink, M=M, B=B)
    if verbose:
        print(">Step 1, Iteration 2", file=sink)
    # iteration 2 feeds the iteration-1 structure back as restraints for every sequence
    pred_dbn = SQRNdbnali(objs, pred_dbn, defreacts, defref, bpweights, interchainonly, minlen, minbpscore,
                          threads, verbose, sink=sink, M=M, B=B)[0]
    step1dbn = PairsToDBN(DBNToPairs(pred_dbn), N, levellimit=levellimit)
    smat = smat / np.max(smat) * 5                                      # ali.py:371
    if verbose:
        print(">Step 1, Result", file=sink)
        print(step1dbn, file=sink)

    if step3 != '1':
        if verbose:
            print(">Step 2, Individuals", file=sink)
        buf = io.StringIO()
        results = _seq.RunSQRNdbnseqBatch([tuple(obj) for obj in objs], paramsetnames, paramsets, rankbydiff,
                                          rankby, hardrest, interchainonly, toplim, outplim, conslim, reactformat,
                                          False, poollim, sink=buf, stemmatrix=smat, algos=algos, entropy=entropy,
                                          M=M, B=B)
        if verbose:
            print(buf.getvalue(), end='', file=sink)
        structs = [r[0] for r in results]
        step2dbn = Consensus(structs, freqlimit, verbose, sink=sink)
        if verbose:
            print(">Step 2, Consensus", file=sink)
            for lim in range(0, 101, 5):
                print(Consensus(structs, lim / 100), str(lim) + '%', sep='\t', file=sink)
    else:
        step2dbn = '.' * N
    step2dbn = PairsToDBN(DBNToPairs(step2dbn), N, levellimit=levellimit)

    if verbose:
        print("=" * N, file=sink)
    seq0 = objs[0][1]
    if defreacts:
        print(EncodedReactivities(seq0, defreacts, reactformat), "reactivities", sep='\t', file=sink)
    if defrests:
        print(''.join(seq0[k] if seq0[k] in SEPS else defrests[k] for k in range(N)), "restraints", sep='\t', file=sink)
    if defref:
        print(''.join(seq0[k] if seq0[k] in SEPS else defref[k] for k in range(N)), "reference", sep='\t', file=sink)
    if defreacts or defref or defrests:
        print("_" * N, file=sink)

    def react_col(dbn):
        return ('\t' + str(round(ReactScore(defreacts, seq0, dbn), 2))) if defreacts else ''

    def metric_col(dbn):
        return "TP={},FP={},FN={},FS={},PR={},RC={}".format(*Metrics(defref, dbn)) if defref else ''

    print(step1dbn, "Step-1" + react_col(step1dbn), metric_col(step1dbn), sep='\t', file=sink)
    skipped = step3 == '1'
    print(step2dbn, "Step-2" + ("(skipped)" if skipped else "") + ("" if skipped else react_col(step2dbn)),
          "" if skipped else metric_col(step2dbn), sep='\t', file=sink)

    if step3 == '1':
        step3dbn = step1dbn
    elif step3 == '2':
        step3dbn = step2dbn
    elif step3 == 'i':
        step3dbn = PairsToDBN(sorted(set(DBNToPairs(step1dbn)) & set(DBNToPairs(step2dbn))), N)
    else:
        pairs = DBNToPairs(step1dbn)
        taken = {p for bp in pairs for p in bp}
        pairs += [(v, w) for v, w in DBNToPairs(step2dbn) if v not in taken and w not in taken]
        step3dbn = PairsToDBN(sorted(pairs), N)
    print(step3dbn, "Step-3({})".format(step3) + react_col(step3dbn), metric_col(step3dbn), sep='\t', file=sink)

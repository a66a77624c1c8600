"""The device functions of csrc/sqrn_device.cuh compiled for the host with a team of ONE
thread (tests/emu) against the oracle: checks the bit-mask enumeration, the scoring, the
per-stem level restatement and ChooseStems without a GPU.  The parallel behaviour of the
same code is covered by the -m gpu tests."""
import json
import os
import random

import numpy as np
import pytest

from oracle import oracle as O
from squarna_b200 import SQRNdbnseq as S
from tests import common as T
from tests.emu import emu

G = os.path.join(os.path.dirname(__file__), "golden")
_OPEN = "([{<ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_CLOSE = ")]}>abcdefghijklmnopqrstuvwxyz"


# competing helices with runs of more than 32 cells (the persistent list cuts those by a cell walk,
# not by the 32-bit mask) and low-complexity repeats (many long overlapping runs)
LONG_RUNS = ["G" * 40 + "AAAA" + "C" * 40 + "AAAA" + "G" * 40,
             "G" * 36 + "U" + "C" * 45 + "GAAA" + "G" * 38 + "A" + "C" * 20,
             "GC" * 50, "GGGCCC" * 20, "A" * 35 + "GAAA" + "U" * 50 + "GCGC" + "A" * 40]


def _prep_batch(cases):
    preps = [S._prepare(c[0], c[1], c[2], None) for c in cases]
    table, codes = {}, []
    for p in preps:
        codes.append(np.array([table.setdefault(float(x), len(table)) for x in p.shortreacts], np.uint16))
    kw = dict(react_codes=codes, react_values=np.array(list(table.keys())),
              restr_class=[p.rclass for p in preps],
              rbps=[np.array(p.rbps, np.int32).reshape(-1, 2) for p in preps])
    return preps, kw


@pytest.mark.parametrize("region", [0, 1, 2, "fast", "runlist", "persist", "persist-small", "persist-general", "glist", "glist-small",
                                    "base", "base-small", "glist-simple"],
                         ids=["auto", "scan", "stems", "fastflavour", "runlist", "persist", "persist-smalllist", "persist-general",
                              "glist", "glist-smalllist", "base-list", "base-list-overflow", "glist-whole-list-sweeps"])
@pytest.mark.parametrize("ps,ccap", [(T.FASTEST, 128), (T.DEFG1, 128), (T.DEFG2, 16), (T.ALI, 64)],
                         ids=["fastest", "defG1", "defG2-smalllist", "ali"])
def test_tail_plain(ps, ccap, region):
    """single-path greedy (pl=1) on plain sequences: stems, dbn, raw scores; the ScoreStems region
    evaluated by the reference's position scan, by the stem walk, and by the automatic choice"""
    seqs = T.rand_seqs(31, 150, 5, 210) + T.rand_seqs(36, 40, 5, 150, "ACGUN") + LONG_RUNS
    if region == "fast":       # compile-time flavour of the byseq fast lane: bit planes, stem walk only
        r = emu.run(ps, seqs, ccap=ccap, flavour=1)
    elif region == "runlist":  # the warp-team scan (shared run list) of the general flavour
        r = emu.run(ps, seqs, ccap=ccap, flavour=2)
    elif region == "persist":  # what k_fast runs: the run list persists across greedy steps
        before = emu.lib().emu_persist_steps()
        r = emu.run(ps, seqs, ccap=ccap, flavour=3, pcap=4096)
        assert emu.lib().emu_persist_steps() - before == r["ncalls"]        # every OptimalStems pass used the list
    elif region == "persist-small":   # a list that overflows on the longer sequences (at build time or while
        r = emu.run(ps, seqs, ccap=ccap, flavour=3, pcap=96)      # pieces are appended): falls back to rescanning
    elif region == "persist-general":
        r = emu.run(ps, seqs, ccap=ccap, flavour=4, pcap=4096)
    elif region == "glist":    # what CTA teams run: the list in global memory, adjusted scores cached between steps
        r = emu.run(ps, seqs, ccap=ccap, flavour=5, pcap=1 << 16)
    elif region == "glist-small":
        r = emu.run(ps, seqs, ccap=ccap, flavour=5, pcap=600)
    elif region == "glist-simple":    # k_long<8, simple>: the list swept whole every pass (small lists)
        r = emu.run(ps, seqs, ccap=ccap, flavour=8, pcap=1 << 16)
    elif region == "base":     # what the tails of pools run: every step sweeps the sequence's base list
        r = emu.run(ps, seqs, ccap=ccap, flavour=6, pcap=1 << 16)
    elif region == "base-small":      # slots too small for the longer sequences: those enumerate as before
        r = emu.run(ps, seqs, ccap=ccap, flavour=6, pcap=300)
    else:
        r = emu.run(ps, seqs, ccap=ccap, region_mode=region)
    for b, s in enumerate(seqs):
        _, structs, _ = O.predict_short(s, [0.5] * len(s), "." * len(s), [ps], poollim=1)
        dbn, sc, isint, _, stems, _, _ = structs[0]
        o = r["dbn_off"][b]
        assert bytes(r["dbn_ascii"][o:o + len(s)]).decode() == dbn
        assert [tuple(int(x) for x in r["stems"][r["off"][b] + k]) for k in range(r["n"][b])] == stems
        assert tuple(emu.lib().emu_pyround3(float(x)) for x in r["raw"][b]) == sc
        assert bool(r["flags"][b] & 1) == isint


@pytest.mark.parametrize("region,flavour", [(1, 0), (2, 0), (0, 2), (0, 4), (0, 5), (0, 6), (0, 8)],
                         ids=["scan", "stems", "runlist", "persist", "glist", "base-list", "glist-whole-list-sweeps"])
@pytest.mark.parametrize("interchain", [False, True])
def test_tail_with_restraints_and_reactivities(interchain, region, flavour):
    rng = random.Random(32)
    cases = [T.rand_case(rng, 8, 150, p_gap=0.2) for _ in range(200)]
    preps, kw = _prep_batch(cases)
    for comp in (False, True):
        idx = [k for k, p in enumerate(preps) if p.compensated == comp]
        sub = {k: [v[i] for i in idx] if isinstance(v, list) else v for k, v in kw.items()}
        for ps in (T.DEFG1, T.FASTEST):
            r = emu.run(ps, [preps[k].shortseq for k in idx], react_comp=comp, interchainonly=interchain,
                            region_mode=region, flavour=flavour, pcap=4096, **sub)
            for b, k in enumerate(idx):
                p = preps[k]
                _, structs, _ = O.predict_short(p.shortseq, p.shortreacts, p.shortrest, [ps], interchainonly=interchain,
                                                poollim=1, compensated_sum=comp)
                dbn, sc, isint, _, stems, _, _ = structs[0]
                got = [tuple(int(x) for x in r["stems"][r["off"][b] + q]) for q in range(r["n"][b])]
                assert got == stems, (p.shortseq, p.shortrest)
                assert tuple(emu.lib().emu_pyround3(float(x)) for x in r["raw"][b]) == sc, (p.shortseq, cases[k][1])
                codes = r["dbn_code"][r["dbn_off"][b]:r["dbn_off"][b] + len(p.shortseq)]
                asc = "".join("." if c == 0 else (_OPEN[c - 1] if c > 0 else _CLOSE[-c - 1]) for c in codes.tolist())
                assert asc == dbn


def test_yield_matches_annotate():
    rng = random.Random(33)
    cases = [T.rand_case(rng, 8, 200, p_gap=0.0) for _ in range(120)]
    preps, kw = _prep_batch(cases)
    for ps in (T.ALI, T.FASTEST):
        r = emu.run(ps, [p.shortseq for p in preps], mode=emu.MODE_YIELD, **kw)
        for b, p in enumerate(preps):
            want = O.annotate(p.shortseq, ps, p.shortreacts, p.shortrest)
            lo = r["off"][b]
            got = [(int(r["stems"][lo + q][0]), int(r["stems"][lo + q][1]), int(r["stems"][lo + q][2]), float(r["fin"][lo + q]))
                   for q in range(r["n"][b])]
            assert got == want, (p.shortseq, p.shortrest)


def test_step_matches_golden_optimal():
    """MODE_STEP (one OptimalStems + ChooseStems) against the REFERENCE's own outputs"""
    with open(os.path.join(G, "optimal.json")) as f:
        cases = json.load(f)
    for c in cases:
        ps = dict(c["ps"])
        ps["algorithms"] = set(ps["algorithms"])
        r = emu.run(ps, [c["seq"]], mode=emu.MODE_STEP, init_stems=[[tuple(s) for s in c["selected"]]],
                    item_subopt=[c["subopt"]], ccap=4096, stem_cap=256)
        n = r["n"][0]
        got = [[int(r["stems"][q][0]), int(r["stems"][q][1]), int(r["stems"][q][2]), float(r["fin"][q])] for q in range(n)]
        assert got == c["chosen"], c["seq"]


def test_per_stem_levels_equal_per_pair_levels():
    """SURVEY A-7: pseudoknot levels computed per stem == PairsToDBN per pair"""
    rng = random.Random(34)
    for _ in range(300):
        n = rng.randint(30, 200)
        used, stems = set(), []
        for _ in range(rng.randint(1, 14)):
            i, j, ln = rng.randrange(n), rng.randrange(n), rng.randint(1, 7)
            if i > j:
                i, j = j, i
            if j - i < 2 * ln + 2:
                continue
            pos = set(range(i, i + ln)) | set(range(j - ln + 1, j + 1))
            if pos & used:
                continue
            used |= pos
            stems.append((i, j, ln))
        if not stems:
            continue
        seq = "A" * n
        r = emu.run(T.DEFG1, [seq], mode=3, init_stems=[stems])          # MODE_FINAL: levels + dbn of given stems
        codes = r["dbn_code"][:n]
        pairs = [(i + k, j - k) for i, j, ln in stems for k in range(ln)]
        want = O.pair_levels(pairs)
        for (v, w), lev in want.items():
            assert codes[v] == lev and codes[w] == -lev, (stems, v, w)


@pytest.mark.parametrize("region", [1, 2, "base", "base-smalllist"], ids=["scan", "stems", "base-list", "base-list-smalllist"])
def test_step_on_random_pseudoknotted_structures(region):
    """ScoreStems on top of arbitrary (pseudoknotted, multi-level) partial structures: the stem walk
    and the position scan must both reproduce the oracle's ChooseStems list and scores"""
    rng = random.Random(35)
    for ps, subopt in ((T.DEFG1, 0.3), (T.DEFG2, 0.5), (T.FASTEST, 0.2)):
        for _ in range(60):
            n = rng.randint(40, 180)
            seq = T.rand_seq(rng, n, "ACGU" if rng.random() < 0.8 else "ACGU;")
            used, stems = set(), []
            for _ in range(rng.randint(1, 12)):
                i, j, ln = rng.randrange(n), rng.randrange(n), rng.randint(1, 6)
                if i > j:
                    i, j = j, i
                if j - i < 2 * ln + 2:
                    continue
                pos = set(range(i, i + ln)) | set(range(j - ln + 1, j + 1))
                if pos & used or any(seq[p] == ";" for p in pos):
                    continue
                used |= pos
                stems.append((i, j, ln))
            _, chosen = O.optimal(seq, ps, subopt, selected=stems)
            if isinstance(region, str):
                # the pool-round path of CTA teams: the candidates come from the sequence's base list (runs of the empty
                # structure, cut on the fly by this structure's unpaired mask) instead of an enumeration
                before = emu.lib().emu_base_sweeps()
                r = emu.run(ps, [seq], mode=emu.MODE_STEP, init_stems=[stems], item_subopt=[subopt],
                            ccap=4096 if region == "base" else 48, stem_cap=512, flavour=6, pcap=1 << 15)
                assert emu.lib().emu_base_sweeps() > before
                if r["n"][0] < 0:                     # (the in-range candidates alone overflow the small list: the host retries)
                    assert region == "base-smalllist"
                    continue
            else:
                r = emu.run(ps, [seq], mode=emu.MODE_STEP, init_stems=[stems], item_subopt=[subopt], ccap=4096,
                            stem_cap=512, region_mode=region)
            k = r["n"][0]
            got = [(int(r["stems"][q][0]), int(r["stems"][q][1]), int(r["stems"][q][2]), float(r["fin"][q])) for q in range(k)]
            assert got == chosen, (seq, stems)


@pytest.fixture
def rebuild_period(request):
    """rebuild period of the binned global list (passes); restored to the default afterwards"""
    emu.lib().emu_gl_set_rebuild(request.param)
    yield request.param
    emu.lib().emu_gl_set_rebuild(0)


@pytest.mark.parametrize("rebuild_period", [0, 1, 3, 7], indirect=True, ids=["default", "every-pass", "every-3", "every-7"])
def test_glist_cached_scores_under_pseudoknots(rebuild_period):
    """The global persistent list (CTA teams) keeps the adjusted score of a candidate between greedy steps
    unless the new stem meets the window ScoreStems read, or an older stem changed its pseudoknot level.
    Parameter sets that accept many pseudoknots (up to ~10 levels) against the rescanning flavour and,
    for a sample, the oracle.  The list is binned by a static bound: records behind the prefix a pass looks at
    are caught up from the unpaired mask when they enter it, and every `rebuild_period` passes the list is
    compacted and re-binned -- both paths must have run."""
    pk = dict(T.ALI); pk["orderpenalty"] = 0.1; pk["minfinscorefactor"] = 0.8
    before = emu.lib().emu_gl_rebuilds(), emu.lib().emu_gl_catchups()
    for ps in (T.ALI, T.DEFG2, T.G1000, pk):
        seqs = T.rand_seqs(101, 40, 150, 450) + T.rand_seqs(102, 10, 300, 500, "GC") + T.rand_seqs(103, 10, 200, 400, "GGCCAU")
        a = emu.run(ps, seqs, ccap=256, flavour=5, pcap=1 << 20)
        b = emu.run(ps, seqs, ccap=256, flavour=2)
        for k in range(len(seqs)):
            sa = [tuple(int(x) for x in a["stems"][a["off"][k] + q]) for q in range(a["n"][k])]
            sb = [tuple(int(x) for x in b["stems"][b["off"][k] + q]) for q in range(b["n"][k])]
            assert sa == sb and (a["raw"][k] == b["raw"][k]).all(), seqs[k]
            assert bytes(a["dbn_ascii"][a["dbn_off"][k]:a["dbn_off"][k + 1]]) == bytes(b["dbn_ascii"][b["dbn_off"][k]:b["dbn_off"][k + 1]])
        for k in range(0, len(seqs), 12):
            _, structs, _ = O.predict_short(seqs[k], [0.5] * len(seqs[k]), "." * len(seqs[k]), [ps], poollim=1)
            sa = [tuple(int(x) for x in a["stems"][a["off"][k] + q]) for q in range(a["n"][k])]
            assert sa == structs[0][4]
    assert emu.lib().emu_gl_rebuilds() > before[0]
    assert emu.lib().emu_gl_catchups() > before[1] or rebuild_period == 1      # (rebuilt every pass, nothing is ever behind)


@pytest.mark.parametrize("flavour", [5, 2, 55], ids=["global-list", "rescan", "global-list-rebuild-every-5"])
def test_device_code_on_rrna_scale_reference_cases(flavour):
    """the device functions (host emulation) on the reference's OWN results for plain sequences of 2050 .. 2500 nt
    (tests/golden/seq_api_xlong.json): the global candidate list with cached scores, and the rescanning pass"""
    from squarna_b200 import SQUARNA as CLI
    pkg = os.path.dirname(os.path.abspath(CLI.__file__))
    with open(os.path.join(G, "seq_api_xlong.json")) as f:
        cases = json.load(f)
    if flavour == 55:
        emu.lib().emu_gl_set_rebuild(5)
        flavour = 5
    for c in cases:
        if flavour == 2 and c["conf"] != "fastest":
            continue                                  # minlen 2 at 2050 nt: minutes in the single-thread rescanning build
        ps = [p for p in CLI.ParseConfig(os.path.join(pkg, c["conf"] + ".conf"))[1] if p["algorithms"] == {"G"} and not p["bpp"]][0]
        try:
            r = emu.run(ps, [c["seq"]], ccap=4096, flavour=flavour, pcap=1 << 21)
        finally:
            emu.lib().emu_gl_set_rebuild(0) if c is cases[-1] else None
        dbn, sc, _psl = c["structs"][0]
        assert bytes(r["dbn_ascii"][:len(c["seq"])]).decode() == dbn == c["cons"]
        got = tuple(emu.lib().emu_pyround3(float(x)) for x in r["raw"][0])
        assert got == tuple(float(x) for x in sc)


@pytest.mark.parametrize("flavour", [5, 2], ids=["global-list", "rescan"])
def test_device_code_on_long_reference_cases(flavour):
    """the device functions (host emulation) on the reference's own results for 321 .. 1137 nt sequences with
    restraints, reactivities, gaps and separators, single-path greedy (tests/golden/seq_api_long.json, pl = 1)"""
    from squarna_b200 import SQUARNA as CLI
    pkg = os.path.dirname(os.path.abspath(CLI.__file__))
    with open(os.path.join(G, "seq_api_long.json")) as f:
        cases = [c for c in json.load(f) if c["poollim"] == 1 and not c["kw"]["hardrest"]]
    assert len(cases) >= 8
    for c in cases:
        ps = [p for p in CLI.ParseConfig(os.path.join(pkg, c["conf"] + ".conf"))[1] if p["algorithms"] == {"G"} and not p["bpp"]][0]
        preps, kw = _prep_batch([(c["seq"], c["reacts"], c["restraints"])])
        p = preps[0]
        r = emu.run(ps, [p.shortseq], react_comp=p.compensated, interchainonly=c["kw"]["interchainonly"],
                    ccap=4096, flavour=flavour, pcap=1 << 20, **kw)
        short = bytes(r["dbn_ascii"][:len(p.shortseq)]).decode()
        want_dbn, want_sc, _ = c["structs"][0]
        assert S.ReAlign(short, p.seq) == want_dbn == c["cons"], (c["conf"], len(c["seq"]))
        got = tuple(emu.lib().emu_pyround3(float(x)) for x in r["raw"][0])
        assert got == tuple(float(x) for x in want_sc) and bool(r["flags"][0] & 1) == (type(want_sc[1]) is int)


def test_random_parameter_sets():
    """a seeded slice of tests/fuzz_emu.py: random parameter sets (pair weights, minlen 2..5, thresholds, distance / order /
    loop terms in and beyond the range of the shipped .conf files, small maxstemnum) for every list flavour, run to
    completion, with restraints / reactivities / interchainonly, single OptimalStems passes on partial structures, and
    150-500 nt through the global candidate list with random rebuild periods"""
    from tests import fuzz_emu
    assert fuzz_emu.campaign(20261018, 4) is None
    assert fuzz_emu.campaign_extras(20261018, 2) is None
    assert fuzz_emu.campaign_long(20261018, 2) is None

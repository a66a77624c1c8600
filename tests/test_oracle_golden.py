"""The CPU oracle (oracle/sqrn_oracle.c) against golden vectors produced by the real
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def _ps(d):
    d = dict(d)
    d["algorithms"] = set(d["algorithms"])
    return d


@pytest.fixture(scope="module")
def configs():
    return load("configs.json")


def _gsets(configs, conf):
    return [_ps(p) for p in configs[conf]["paramsets"] if p["algorithms"] == ["G"] and p["bpp"] == 0]


@pytest.mark.parametrize("fname,least", [("seq_api.json", 250), ("seq_api_long.json", 20), ("seq_api_xlong.json", 1),
                                         ("seq_api_c3.json", 5), ("seq_api_c3b.json", 7)],
                         ids=["short", "321-1200nt", "2050-2500nt", "config3-shape", "config3-550-1500nt"])
def test_seq_api(configs, fname, least):
    """SQRNdbnseq end to end: consensus, every structure, its three scores (incl. the int-0
    quirk) and its parameter-set list, in rank order.  The second file holds sequences of 321 .. 1137 nt,
    the lengths the CTA-team kernels serve; the third three plain sequences of 2050 .. 2500 nt (1024-thread CTAs and
    clusters; minutes each in the reference); the fourth the shape of BASELINE config 3 (300 .. 620 nt, reactivity
    letters, restraints incl. planted stems, G sets by length, pl=100: up to 173 ranked structures per sequence); the
    fifth eight config-3 inputs from workloads.config3 at 550 .. 1500 nt (500nobpp's two G sets below 1000 nt, 1000nobpp above)."""
    cases = load(fname)
    assert len(cases) > least
    for c in cases:
        kw = dict(c["kw"])
        kw["rankby"] = tuple(kw["rankby"])
        smat = None if c["smat"] is None else np.array(c["smat"])
        cons, structs = O.sqrn_dbnseq(c["seq"], c["reacts"], c["restraints"], paramsets=_gsets(configs, c["conf"]),
                                      poollim=c["poollim"], stemmatrix=smat, **kw)
        assert cons == c["cons"], c["seq"]
        assert len(structs) == len(c["structs"]), c["seq"]
        for (d, sc, psl), (gd, gsc, gpsl) in zip(structs, c["structs"]):
            assert d == gd and list(sc) == gsc and psl == gpsl, (c["seq"], d, gd, sc, gsc)
            assert type(sc[1]) is type(gsc[1])


def test_seq_api_random_parameter_sets():
    """the same end-to-end comparison under RANDOM parameter sets (make_golden.py fuzz: pair weights, minlen 2..5,
    thresholds, distance / order / loop terms in and beyond the range of the shipped .conf files, small maxstemnum, 1-3
    sets per call with random subopt ranges, random poollim) -- 1000 committed cases; the generating runs compared 20 600 more"""
    cases = load("seq_api_fuzz.json")
    assert len(cases) >= 1000 and sum(len(c["paramsets"]) > 1 for c in cases) > 50 and sum(c["poollim"] > 1 for c in cases) > 100
    for c in cases:
        kw = dict(c["kw"])
        kw["rankby"] = tuple(kw["rankby"])
        cons, structs = O.sqrn_dbnseq(c["seq"], c["reacts"], c["restraints"], paramsets=[_ps(p) for p in c["paramsets"]],
                                      poollim=c["poollim"], **kw)
        assert cons == c["cons"], c["seq"]
        assert len(structs) == len(c["structs"]), c["seq"]
        for (d, sc, psl), (gd, gsc, gpsl) in zip(structs, c["structs"]):
            assert d == gd and list(sc) == gsc and psl == gpsl, (c["seq"], d, gd, sc, gsc)
            assert type(sc[1]) is type(gsc[1])


def test_annotate_stems():
    """BPMatrix + AnnotateStems: same stems, same order, same float64 scores"""
    for c in load("annotate.json"):
        got = O.annotate(c["seq"], _ps(c["ps"]), c["reacts"], c["restraints"], interchainonly=c["interchainonly"])
        assert [list(x) for x in got] == c["stems"], c["seq"]


def test_levels():
    for c in load("levels.json"):
        lev = O.pair_levels(c["pairs"])
        assert sorted([k[0], k[1], v] for k, v in lev.items()) == c["levels"]


def test_optimal_stems():
    """ScoreStems survivors (order + adjusted scores) and the ChooseStems list"""
    for c in load("optimal.json"):
        scored, chosen = O.optimal(c["seq"], _ps(c["ps"]), c["subopt"], selected=[tuple(s) for s in c["selected"]])
        assert [list(x) for x in scored] == c["scored"], c["seq"]
        assert [list(x) for x in chosen] == c["chosen"], c["seq"]

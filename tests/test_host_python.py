"""Host-side Python of squarna_b200 (parsers, config, dbn helpers, reactivity handling,
text output) against golden vectors produced by the real reference.  No GPU needed."""
import contextlib
import io
import json
import math
import os

import pytest

from squarna_b200 import SQRNdbnseq as S
from squarna_b200 import SQUARNA as CLI

G = os.path.join(os.path.dirname(__file__), "golden")
PKG = os.path.dirname(os.path.abspath(CLI.__file__))


def load(name):
    with open(os.path.join(G, name)) as f:
        return json.load(f)


def test_shipped_configs_parse_like_the_reference():
    """every .conf shipped in the package gives the names and values the reference's files give"""
    gold = load("configs.json")
    assert len(gold) == 17
    for conf, g in gold.items():
        names, psets = CLI.ParseConfig(os.path.join(PKG, conf + ".conf"))
        assert names == g["names"], conf
        assert len(psets) == len(g["paramsets"])
        for ps, gp in zip(psets, g["paramsets"]):
            ps = dict(ps)
            ps["algorithms"] = sorted(ps["algorithms"])
            assert ps == gp, (conf, ps, gp)
            assert list(ps["bpweights"]) == list(gp["bpweights"])      # dict order matters (seq.py:282-284)


def test_config_errors(tmp_path):
    p = tmp_path / "bad.conf"
    p.write_text(">x\nalgorithms G\nbpweights GC=1\n")
    with pytest.raises(ValueError, match="Missing some of the parameters"):
        CLI.ParseConfig(str(p))


def test_reactivity_helpers():
    for c in load("reacts.json"):
        vals = [float("nan") if v is None else v for v in c["vals"]]
        out = S.ProcessReacts(vals, M=c["M"], B=c["B"])
        assert [float(x) for x in out] == c["out"]
        for f, enc in c["enc"].items():
            assert S.EncodedReactivities(c["seq"], out, int(f)) == enc
    assert S.ProcessReacts([]) == []
    assert S.ReactDict["z"] == 1.0 and S.ReactDict["d"] == 0.12 and S.ReactDict["?"] == -999


def test_pairs_to_dbn_and_levels():
    for c in load("levels.json"):
        pairs = [tuple(p) for p in c["pairs"]]
        assert S.PairsToDBN(pairs, c["n"]) == c["dbn"]
        lev = S.PairsToDBN(pairs, returnlevels=True)
        assert sorted([k[0], k[1], v] for k, v in lev.items()) == c["levels"]
        assert set(S.DBNToPairs(c["dbn"])) == set(pairs) or max(v for _, _, v in c["levels"]) > 49


def test_dbn_helpers():
    assert S.DBNToPairs("((..[[..))..]]") == [(0, 9), (1, 8), (4, 13), (5, 12)]
    assert S.DBNToPairs("))((") == []                       # closers without openers are ignored
    assert S.DBNToPairs("Бб") == [(0, 1)]
    assert S.UnAlign("AC-GU", "(.(.)") == ("ACGU", "(...")   # the pair (2,4) touches a gap column and vanishes
    assert S.UnAlign("A-CGU", "(..).") == ("ACGU", "(.).")
    assert S.ReAlign("(.).", "A-CGU") == "(..)."
    assert S.ReAlign("ACGU", "A-CGU", seqmode=True) == "A-CGU"
    with pytest.raises(AssertionError, match="Cannot ReAlign"):
        S.ReAlign("(.)", "A-CGU")
    assert S.PairsToStems([(0, 9), (1, 8), (4, 13)]) == [[[(0, 9), (1, 8)], 2], [[(4, 13)], 1]]
    assert S.ParseRestraints("(_/\\+)") == ([(0, 5)], {1, 4}, {2}, {3})


def test_default_format_parser():
    path = os.path.join(G, "inputs", "seq_input.fas")
    entries = list(CLI.ParseDefaultInput(path, "qtrf"))
    assert len(entries) == 17
    by_name = {e[0]: e for e in entries}
    name, seq, reacts, rest, ref = by_name[">External loop 1"]
    assert seq == "CCCAAAAGGG;CCCAAAAGGG" and reacts is None       # the trailing comment is dropped
    assert len(by_name[">testcase with reactivities"][2]) == 16
    assert by_name[">multiple chains with bp-to-the-right restraints"][3] == "........./////....................."
    assert CLI.GuessFormat(path) == ("default", False)
    assert CLI.GuessFormat(os.path.join(G, "inputs", "ali_input.afa"))[0] == "default"
    d = next(CLI.ParseDefaultInput(os.path.join(G, "inputs", "ali_input.afa"), "qtrf", returndefaults=True))
    assert d[0] is not None and d[2] is not None and len(d[0]) == len(d[2]) == 115


def test_other_parsers(tmp_path):
    fa = tmp_path / "x.fa"
    fa.write_text(">a\nACGU\nACGU\n\n>b\nGGGG\n")
    assert list(CLI.ParseFasta(str(fa))) == [(">a", "ACGUACGU", None, None, None), (">b", "GGGG", None, None, None)]
    assert CLI.GuessFormat(str(fa)) == ("fasta", False)
    stk = tmp_path / "x.stk"
    stk.write_text("# STOCKHOLM 1.0\n#=GF ID test\ns1 ACGU\ns2 AC-U\n#=GC SS_cons (..)\n\ns1 GG\ns2 GG\n#=GC SS_cons ..\n//\n")
    objs, single = CLI.ParseStockholm(str(stk))
    assert objs == [(">s1", "ACGUGG", None, None, "(..)..",), (">s2", "AC-UGG", None, None, "(..)..")] and not single
    assert CLI.ParseStockholm(str(stk), True) == (None, None, "(..)..")
    aln = tmp_path / "x.aln"
    aln.write_text("CLUSTAL W\n\ns1 ACGU\ns2 AC-U\n   **\n\ns1 GG\ns2 GG\n")
    assert CLI.ParseClustal(str(aln)) == ([(">s1", "ACGUGG", None, None, None), (">s2", "AC-UGG", None, None, None)], False)
    assert CLI.GuessFormat(str(aln)) == ("clustal", 0) and CLI.GuessFormat(str(stk)) == ("stockholm", 0)


def test_predict_argument_errors():
    with pytest.raises(AssertionError, match="Input file does not exist"):
        CLI.Predict(inputfile="/nonexistent")
    with pytest.raises(AssertionError, match="Config file does not exist"):
        CLI.Predict(inputseq="ACGU", configfile="nope")
    with pytest.raises(ValueError, match="Inappropriate toplim"):
        CLI.Predict(inputseq="ACGU", configfile="fastest", toplim="x")
    with pytest.raises(AssertionError, match="Inappropriate rankby"):
        CLI.Predict(inputseq="ACGU", configfile="fastest", rankby="q")
    with pytest.raises(ValueError, match="Inappropriate algorithm"):
        CLI.Predict(inputseq="ACGU", configfile="fastest", algorithms="x")
    with pytest.raises(NotImplementedError):
        CLI.Predict(inputseq="ACGU", configfile="fastest", entropy=True)


def test_evalonly_cli_text_matches_reference():
    """`eo` needs no prediction: parser + ReferenceScores + text are host code (byte-equal output)"""
    argv = load("cli_manifest.json")["seq_evalonly"]
    buf = io.StringIO()
    cwd = os.getcwd()
    os.chdir(G)
    try:
        with contextlib.redirect_stdout(buf):
            CLI.Main(argv)
    finally:
        os.chdir(cwd)
    with open(os.path.join(G, "cli", "seq_evalonly.txt")) as f:
        assert buf.getvalue() == f.read()

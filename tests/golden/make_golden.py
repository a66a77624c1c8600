#!/usr/bin/env python
"""Generates the golden vectors in this directory by IMPORTING THE REAL REFERENCE
(/root/reference/src/SQUARNA, febos/SQUARNA) in the build container.

    python tests/golden/make_golden.py          # JSON vectors
    python tests/golden/make_golden.py cli      # CLI text of the reference on tests/golden/inputs/
    python tests/golden/make_golden.py algos    # SQRNdbnseq with the Nussinov / Hungarian / Edmonds parameter sets
    python tests/golden/make_golden.py long     # SQRNdbnseq on 321 .. 1200 nt sequences (minutes)
    python tests/golden/make_golden.py xlong    # three sequences of 2050 .. 2500 nt (tens of minutes)
    python tests/golden/make_golden.py c3       # BASELINE config 3 shape (reactivities + restraints, pl=100), 300 .. 620 nt
    python tests/golden/make_golden.py c3b N    # one config-3 case of N nt from workloads.config3 (700 / 1100 / 1500: tens of minutes each)
    python tests/golden/make_golden.py xlong3k  # one plain 3000-nt sequence, 1000nobpp G, pl=1 (config 5 lengths)
    python tests/golden/make_golden.py c4       # config 4 shape: the reference CLI on a synthetic 64 x 400 alignment

The reference cannot travel to the GPU box, so its outputs are committed here as
JSON fixtures; tests/test_oracle_golden.py pins the CPU oracle (and the host-side
Python of squarna_b200) to them.  Python 3.12.3 / numpy 2.3.5 / glibc 2.39 were
used (float repr round-trips exactly through JSON).
"""
import io
import json
import os
import random
import sys

REF = "/root/reference/src/SQUARNA"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
import SQRNdbnseq as R  # noqa: E402
import SQUARNA as RC  # noqa: E402

from tests.common import rand_case, rand_seq  # noqa: E402

CONFS = ["fastest", "greedynobpp", "nobpp", "500nobpp", "1000nobpp", "alt", "ali",
         "def", "500", "1000", "greedy", "edmondsnobpp", "hungariannobpp", "nussinovnobpp",
         "edmonds", "hungarian", "nussinov"]


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print(name, os.path.getsize(path), "bytes")


def jsonable_ps(ps):
    d = dict(ps)
    d["algorithms"] = sorted(d["algorithms"])
    return d


def gsets(conf):
    names, psets = RC.ParseConfig(os.path.join(REF, conf + ".conf"))
    return [p for p in psets if p["algorithms"] == {"G"} and p["bpp"] == 0]


def main():
    # --- G6: parsed configs ------------------------------------------------
    confs = {}
    for c in CONFS:
        names, psets = RC.ParseConfig(os.path.join(REF, c + ".conf"))
        confs[c] = {"names": names, "paramsets": [jsonable_ps(p) for p in psets]}
    dump("configs.json", confs)

    # --- G1: SQRNdbnseq end to end -------------------------------------------
    rng = random.Random(20261017)
    plan = [("fastest", 1, 8, 160, 90), ("fastest", 1000, 8, 120, 30), ("greedynobpp", 100, 8, 80, 60),
            ("greedynobpp", 5, 8, 110, 40), ("greedynobpp", 1, 8, 140, 30), ("alt", 50, 20, 70, 25),
            ("ali", 1000, 10, 90, 30), ("1000nobpp", 100, 100, 220, 8), ("500nobpp", 100, 60, 120, 6)]
    cases = []
    for conf, pl, lo, hi, count in plan:
        psets = gsets(conf)
        for _ in range(count):
            seq, reacts, rest, kw = rand_case(rng, lo, hi)
            if rng.random() < 0.25 and len(psets) > 1:
                kw["priority"] = [rng.randrange(len(psets))]
            smat = None
            if rng.random() < 0.12 and len(seq) <= 60:
                n = len(seq)
                m = np.array([[rng.randrange(0, 21) / 4.0 for _ in range(n)] for _ in range(n)])
                smat = ((m + m.T) / 2).tolist()
            try:
                out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=pl, algos={"G"},
                                   stemmatrix=None if smat is None else np.array(smat),
                                   **{k: (set(v) if k == "priority" else v) for k, v in kw.items()})
            except ZeroDivisionError:
                continue
            cases.append({"conf": conf, "poollim": pl, "seq": seq, "reacts": reacts, "restraints": rest,
                          "kw": kw, "smat": smat, "cons": out[0],
                          "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
    dump("seq_api.json", cases)

    # --- G2: BPMatrix + AnnotateStems ----------------------------------------
    ann = []
    for conf in ("fastest", "greedynobpp", "ali"):
        for ps in gsets(conf)[:2]:
            for _ in range(25):
                seq, reacts, rest, kw = rand_case(rng, 8, 90, p_gap=0.0)
                seq = seq.upper().replace("T", "U")
                if isinstance(reacts, str):
                    reacts = R.ProcessReacts([R.ReactDict[c] for c in reacts])
                rbps, rxs, rl, rr = R.ParseRestraints(rest or "." * len(seq))
                bm, sm = R.BPMatrix(seq, ps["bpweights"], rxs, rl, rr, kw["interchainonly"], reacts=reacts)
                stems = R.AnnotateStems(bm, sm, rbps, [], ps["minlen"], ps["minbpscore"])
                ann.append({"ps": jsonable_ps(ps), "seq": seq, "reacts": None if reacts is None else [float(x) for x in reacts],
                            "restraints": rest, "interchainonly": kw["interchainonly"],
                            "stems": [[s[0][0][0], s[0][0][1], s[1], float(s[2])] for s in stems]})
    dump("annotate.json", ann)

    # --- G3: PairsToDBN levels / glyphs on random (pseudoknotted) stem sets ----
    lev = []
    for _ in range(120):
        n = rng.randint(20, 120)
        used, pairs = set(), []
        for _ in range(rng.randint(1, 10)):
            i, j, ln = rng.randrange(n), rng.randrange(n), rng.randint(1, 6)
            if i > j:
                i, j = j, i
            st = [(i + k, j - k) for k in range(ln) if i + k < j - k]
            if st and not any(v in used or w in used for v, w in st):
                pairs += st
                used |= {p for bp in st for p in bp}
        levels = R.PairsToDBN(pairs, returnlevels=True)
        lev.append({"n": n, "pairs": pairs, "levels": [[k[0], k[1], v] for k, v in sorted(levels.items())],
                    "dbn": R.PairsToDBN(pairs, n)})
    dump("levels.json", lev)

    # --- G4: OptimalStems on partial structures ---------------------------------
    opt = []
    for conf, subopt in (("fastest", 1.0), ("greedynobpp", 0.65), ("greedynobpp", 0.9), ("ali", 1.0)):
        for ps in gsets(conf)[:2]:
            for _ in range(20):
                seq = rand_seq(rng, rng.randint(30, 130))
                reacts = [0.5] * len(seq)
                bm, sm = R.BPMatrix(seq, ps["bpweights"], set(), set(), set(), False, reacts=reacts)
                # a partial structure: a random prefix of the single-path greedy result
                stems, cur = [], []
                while True:
                    new = R.OptimalStems(seq, cur, bm, sm, reacts, set(), 1.0, ps["minlen"], ps["minbpscore"],
                                         ps["minbpscore"] * ps["minfinscorefactor"], ps["bracketweight"],
                                         ps["distcoef"], ps["orderpenalty"], ps["loopbonus"])
                    if not new:
                        break
                    cur = cur + [new[0]]
                sel = cur[:rng.randint(0, len(cur))]
                allst = R.AnnotateStems(bm, sm, set(), sel, ps["minlen"], ps["minbpscore"])
                scored = R.ScoreStems(seq, allst, sel, reacts, ps["minbpscore"] * ps["minfinscorefactor"],
                                      ps["bracketweight"], ps["distcoef"], ps["orderpenalty"], ps["loopbonus"])
                chosen = R.ChooseStems(scored, subopt)
                opt.append({"ps": jsonable_ps(ps), "subopt": subopt, "seq": seq,
                            "selected": [[s[0][0][0], s[0][0][1], s[1]] for s in sel],
                            "scored": [[s[0][0][0], s[0][0][1], s[1], float(s[2]), float(s[3])] for s in scored],
                            "chosen": [[s[0][0][0], s[0][0][1], s[1], float(s[3])] for s in chosen]})
    dump("optimal.json", opt)

    # --- G7: reactivity helpers -----------------------------------------------
    rx = []
    for _ in range(40):
        vals = [rng.choice([rng.random() * 2 - 0.3, -999.0, float("nan"), 0.0, 1.0, 0.5]) for _ in range(rng.randint(1, 30))]
        seq = rand_seq(rng, len(vals), "ACGU;")
        for M, B in ((1.8, -0.6), (1.8, 1.6)):
            pr = R.ProcessReacts(vals, M=M, B=B)
            rx.append({"vals": [None if v != v else v for v in vals], "M": M, "B": B, "out": [float(x) for x in pr],
                       "seq": seq, "enc": {str(f): R.EncodedReactivities(seq, pr, f) for f in (3, 10, 26)}})
    dump("reacts.json", rx)


def algos_golden():
    """SQRNdbnseq with parameter sets that name Nussinov / Hungarian / Edmonds (bpp 0): nobpp.conf (defG1, defG2,
    defN, defE, defH) and the single-algorithm configs.  Every set names ONE algorithm, so the reference's
    iteration over the `algos` set is deterministic."""
    rng = random.Random(20261018)
    cases = []
    plan = [("nobpp", 100, 8, 90, 40), ("nobpp", 1, 20, 140, 15), ("nussinovnobpp", 100, 8, 120, 25),
            ("edmondsnobpp", 100, 8, 120, 25), ("hungariannobpp", 100, 8, 120, 25)]
    for conf, pl, lo, hi, count in plan:
        names, psets = RC.ParseConfig(os.path.join(REF, conf + ".conf"))
        assert all(p["bpp"] == 0 and len(p["algorithms"]) == 1 for p in psets), conf
        for _ in range(count):
            seq, reacts, rest, kw = rand_case(rng, lo, hi)
            if rng.random() < 0.3:
                kw["priority"] = [rng.randrange(len(psets))]
            if rng.random() < 0.3:
                kw["conslim"] = rng.choice([1, 2, 3])
            if rng.random() < 0.2:
                kw["levellimit"] = rng.choice([1, 2])
            try:
                out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=pl,
                                   **{k: (set(v) if k == "priority" else v) for k, v in kw.items()})
            except ZeroDivisionError:
                continue
            cases.append({"conf": conf, "poollim": pl, "seq": seq, "reacts": reacts, "restraints": rest, "kw": kw,
                          "cons": out[0], "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
    dump("algos.json", cases)
    # SQRNdbnseq(entropy=True): the stem-matrix entropy string of the first parameter set
    ent = []
    for conf in ("greedynobpp", "fastest", "ali"):
        names, psets = RC.ParseConfig(os.path.join(REF, conf + ".conf"))
        for _ in range(12):
            seq, reacts, rest, kw = rand_case(rng, 8, 120)
            out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, entropy=True, interchainonly=kw["interchainonly"])
            ent.append({"conf": conf, "seq": seq, "reacts": reacts, "restraints": rest,
                        "interchainonly": kw["interchainonly"], "entropy": out})
    dump("entropy.json", ent)
    # the same with an alignment-derived stem matrix (step 2 of the alignment mode, seq.py:1031-1034, 1084-1085)
    smat_cases = []
    names, psets = RC.ParseConfig(os.path.join(REF, "nobpp.conf"))
    for _ in range(30):
        seq, reacts, rest, kw = rand_case(rng, 8, 60)
        n = len(seq)
        m = np.array([[rng.randrange(0, 21) / 4.0 for _ in range(n)] for _ in range(n)])
        smat = ((m + m.T) / 2).tolist()
        try:
            out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=100, stemmatrix=np.array(smat), **kw)
        except ZeroDivisionError:
            continue
        smat_cases.append({"conf": "nobpp", "poollim": 100, "seq": seq, "reacts": reacts, "restraints": rest, "kw": kw,
                           "smat": smat, "cons": out[0], "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
    dump("algos_smat.json", smat_cases)


def long_golden():
    """SQRNdbnseq end to end on sequences of 321 .. 1200 nt (the lengths the CTA-team kernels serve): pins the
    oracle, and through it the kernels, where the reference needs seconds to minutes per case"""
    import time
    rng = random.Random(20261020)
    plan = [("fastest", 1, 330, 520, 6), ("1000nobpp", 1, 330, 420, 3), ("greedynobpp", 3, 321, 360, 2),
            ("fastest", 1, 600, 1200, 5), ("1000nobpp", 1, 500, 900, 4), ("500nobpp", 3, 500, 620, 2),
            ("greedynobpp", 5, 330, 420, 2)]
    cases = []
    for conf, pl, lo, hi, count in plan:
        psets = gsets(conf)
        for _ in range(count):
            seq, reacts, rest, kw = rand_case(rng, lo, hi)
            t0 = time.time()
            out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=pl, algos={"G"}, **kw)
            print(conf, len(seq), "%.1f s" % (time.time() - t0), flush=True)
            cases.append({"conf": conf, "poollim": pl, "seq": seq, "reacts": reacts, "restraints": rest, "kw": kw,
                          "smat": None, "cons": out[0], "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
    dump("seq_api_long.json", cases)


def xlong_golden():
    """three plain sequences of 2050 .. 2500 nt (the 1024-thread CTA / cluster kernels): tens of minutes in the reference"""
    import time
    rng = random.Random(20261021)
    cases = []
    for conf, n in (("fastest", 2100), ("fastest", 2500), ("1000nobpp", 2050)):
        psets = gsets(conf)
        seq = rand_seq(rng, n)
        t0 = time.time()
        out = R.SQRNdbnseq(seq, None, None, None, psets, mp=False, poollim=1, algos={"G"})
        print(conf, n, "%.0f s" % (time.time() - t0), flush=True)
        cases.append({"conf": conf, "poollim": 1, "seq": seq, "reacts": None, "restraints": None,
                      "kw": {"rankby": [0, 2, 1]}, "smat": None, "cons": out[0],
                      "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
        dump("seq_api_xlong.json", cases)


def c3_golden():
    """BASELINE config 3 shape: 300 .. 620 nt, reactivities as rf=26 letters (3 % '?'), restraints 5 % '_', 1 % '/',
    1 % '\\' plus 0-2 canonical stems as brackets, G sets by length (greedynobpp < 500, 500nobpp above), CLI default pl=100"""
    import time
    rng = random.Random(20261022)
    comp = {"A": "U", "U": "A", "G": "C", "C": "G"}
    cases = []
    for n in (300, 340, 420, 480, 520, 620):
        seq = [rng.choice("ACGU") for _ in range(n)]
        rest = ["."] * n
        for _ in range(rng.randint(0, 2)):                      # a planted canonical stem, given as a restraint
            ln = rng.randint(3, 6)
            i = rng.randrange(5, n // 2 - 10)
            j = rng.randrange(n // 2 + 10, n - 5)
            for k in range(ln):
                seq[j - k] = comp[seq[i + k]]
                rest[i + k], rest[j - k] = "(", ")"
        for k in range(n):
            if rest[k] == ".":
                x = rng.random()
                rest[k] = "_" if x < 0.05 else "/" if x < 0.06 else "\\" if x < 0.07 else "."
        seq, rest = "".join(seq), "".join(rest)
        reacts = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") if rng.random() > 0.03 else "?" for _ in range(n))
        conf = "greedynobpp" if n < 500 else "500nobpp"
        psets = gsets(conf)
        t0 = time.time()
        out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=100, algos={"G"})
        print(conf, n, "%.0f s" % (time.time() - t0), len(out[1]), "structures", flush=True)
        cases.append({"conf": conf, "poollim": 100, "seq": seq, "reacts": reacts, "restraints": rest,
                      "kw": {"rankby": [0, 2, 1]}, "smat": None, "cons": out[0],
                      "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
        dump("seq_api_c3.json", cases)



def xlong3k_golden():
    """one plain 3000-nt sequence (config 5 lengths), 1000nobpp G set, pl=1: the reference takes tens of minutes"""
    import time
    import workloads
    sym, off, lens = workloads.config5(1, seed=20261023, lo=3000, hi=3000)
    seq = sym.tobytes().decode()
    psets = gsets("1000nobpp")
    t0 = time.time()
    out = R.SQRNdbnseq(seq, None, None, None, psets, mp=False, poollim=1, algos={"G"})
    print("1000nobpp", len(seq), "%.0f s" % (time.time() - t0), flush=True)
    dump("seq_api_xlong3k.json", [{"conf": "1000nobpp", "poollim": 1, "seq": seq, "reacts": None, "restraints": None,
                                   "kw": {"rankby": [0, 2, 1]}, "smat": None, "cons": out[0],
                                   "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]}])


def c3b_golden(n):
    """BASELINE config 3 above 620 nt: ONE case of n nt made by workloads.config3 (reactivity letters, restraints,
    planted stems), G sets by length (500nobpp's two G sets for 500-999 nt, 1000nobpp above), CLI default pl=100"""
    import time
    import workloads
    (seq, reacts, rest), = workloads.config3(1, seed=20261024 + n, lo=n, hi=n)
    conf = workloads.config3_conf(n)
    psets = gsets(conf)
    t0 = time.time()
    out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=100, algos={"G"})
    print(conf, n, "%.0f s" % (time.time() - t0), len(out[1]), "structures", flush=True)
    # (the per-length files are merged into seq_api_c3b.json afterwards)
    dump("seq_api_c3b_%d.json" % n, [{"conf": conf, "poollim": 100, "seq": seq, "reacts": reacts, "restraints": rest,
                                      "kw": {"rankby": [0, 2, 1]}, "smat": None, "cons": out[0],
                                      "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]}])


def c4_golden(n_seqs=64, anc_len=300, n_cols=400, verbose=False):
    """BASELINE config 4 shape (n_seqs x 400): the reference CLI in alignment mode on workloads.config4"""
    import subprocess
    import time
    import workloads
    rows, ref = workloads.config4(n_seqs, anc_len, n_cols)
    name = "ali_c4_%dx%d" % (n_seqs, n_cols)
    with open(os.path.join(HERE, "inputs", name + ".afa"), "w") as f:
        f.write(workloads.config4_text(rows, ref))
    t0 = time.time()
    argv = ["i=inputs/%s.afa" % name, "a"] + (["v"] if verbose else []) + ["t=3"]
    out = subprocess.run([sys.executable, os.path.join(REF, "SQUARNA.py")] + argv, cwd=HERE, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    print(name, "%.0f s" % (time.time() - t0), len(out.stdout), "chars", flush=True)
    with open(os.path.join(HERE, "cli", name + ("_verbose" if verbose else "") + ".txt"), "w") as f:
        f.write(out.stdout)


def bpp_golden():
    """SQRNdbnseq with the bpp != 0 parameter sets (def.conf, greedy.conf, edmonds / hungarian / nussinov.conf): the REAL
    reference with tests/fake_rna.py in place of ViennaRNA's RNA module (not installed here) -- pins the additive /
    multiplicative weighting of the score matrix, the rescale fall-back, the SHAPE hand-over and everything downstream"""
    from tests import fake_rna
    sys.modules["RNA"] = fake_rna
    rng = random.Random(20261025)
    cases = []
    plan = [("def", 100, 8, 70, 30), ("greedy", 100, 8, 90, 25), ("greedy", 1, 30, 160, 10), ("edmonds", 100, 8, 90, 8),
            ("hungarian", 100, 8, 90, 8), ("nussinov", 100, 8, 90, 8), ("def", 5, 20, 100, 8)]
    for conf, pl, lo, hi, count in plan:
        names, psets = RC.ParseConfig(os.path.join(REF, conf + ".conf"))
        assert all(len(p["algorithms"]) == 1 for p in psets), conf
        for _ in range(count):
            seq, reacts, rest, kw = rand_case(rng, lo, hi)
            if conf == "def" and rng.random() < 0.6:
                kw["priority"] = [k for k, nm in enumerate(names) if nm in ("bppN", "bppH1", "bppH2")]     # the CLI's default
            if rng.random() < 0.3:
                kw["conslim"] = rng.choice([1, 2, 3])
            M, B = rng.choice([(1.8, -0.6), (1.8, -0.6), (2.6, -0.8)])
            try:
                out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=pl, M=M, B=B,
                                   **{k: (set(v) if k == "priority" else v) for k, v in kw.items()})
            except ZeroDivisionError:
                continue
            cases.append({"conf": conf, "poollim": pl, "seq": seq, "reacts": reacts, "restraints": rest, "kw": kw, "M": M, "B": B,
                          "cons": out[0], "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
    # one case above 500 nt with 500.conf's sets (autoconfig length class)
    names, psets = RC.ParseConfig(os.path.join(REF, "500.conf"))
    seq, reacts, rest, kw = rand_case(rng, 505, 520)
    out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=3, **kw)
    cases.append({"conf": "500", "poollim": 3, "seq": seq, "reacts": reacts, "restraints": rest, "kw": kw, "M": 1.8, "B": -0.6,
                  "cons": out[0], "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]})
    dump("seq_api_bpp.json", cases)
    # entropy of a first parameter set with bpp (greedy.conf: bppG1)
    ent = []
    names, psets = RC.ParseConfig(os.path.join(REF, "greedy.conf"))
    for _ in range(10):
        seq, reacts, rest, kw = rand_case(rng, 8, 100)
        out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, entropy=True, interchainonly=kw["interchainonly"])
        ent.append({"conf": "greedy", "seq": seq, "reacts": reacts, "restraints": rest,
                    "interchainonly": kw["interchainonly"], "entropy": out})
    dump("entropy_bpp.json", ent)


def fuzz_golden(n_cases, n_keep, seed=20261018, lo=8, hi=110):
    """the reference under RANDOM parameter sets (tests/fuzz_emu.rand_paramset: pair weights, minlen 2..5, thresholds,
    distance / order / loop terms in and beyond the range of the shipped .conf files, small maxstemnum; 1-3 sets per
    call with random subopt ranges) on rand_case inputs (gaps, separators, restraints, reactivities), random poollim:
    every case is compared with the oracle HERE, the first n_keep are written to seq_api_fuzz.json"""
    import time
    from oracle import oracle as O
    from tests.fuzz_emu import rand_paramset
    rng = random.Random(seed)
    cases, bad, t0 = [], 0, time.time()
    while len(cases) < n_cases:
        psets = []
        for _ in range(rng.choice([1, 1, 2, 3])):
            ps = rand_paramset(rng)
            ps["suboptmax"] = rng.choice([1.0, 0.99, 0.95, 0.9, 0.8])
            ps["suboptmin"] = min(ps["suboptmax"], rng.choice([1.0, 0.99, 0.9, 0.65, 0.5]))
            ps["suboptsteps"] = float(rng.choice([1, 1, 2, 3]))
            psets.append(ps)
        pl = rng.choice([1, 1, 3, 20, 100])
        seq, reacts, rest, kw = rand_case(rng, lo, hi)
        if hi > 200:
            pl = min(pl, 3)                          # (the reference needs minutes per pool at these lengths)
        if rng.random() < 0.25 and len(psets) > 1:
            kw["priority"] = [rng.randrange(len(psets))]
        try:
            out = R.SQRNdbnseq(seq, reacts, rest, None, psets, mp=False, poollim=pl, algos={"G"},
                               **{k: (set(v) if k == "priority" else v) for k, v in kw.items()})
        except ZeroDivisionError:
            continue
        case = {"paramsets": [jsonable_ps(p) for p in psets], "poollim": pl, "seq": seq, "reacts": reacts, "restraints": rest,
                "kw": kw, "smat": None, "cons": out[0], "structs": [[d, list(sc), list(ps)] for d, sc, ps in out[1]]}
        cases.append(case)
        okw = dict(kw)
        okw["rankby"] = tuple(okw["rankby"])
        cons, structs = O.sqrn_dbnseq(seq, reacts, rest, paramsets=psets, poollim=pl, **okw)
        same = cons == out[0] and len(structs) == len(out[1]) and all(
            d == gd and list(sc) == list(gsc) and psl == list(gpsl) and type(sc[1]) is type(gsc[1])
            for (d, sc, psl), (gd, gsc, gpsl) in zip(structs, out[1]))
        if not same:
            bad += 1
            print("ORACLE DIFFERS:", json.dumps(case)[:2000], flush=True)
        if len(cases) % (200 if hi <= 200 else 10) == 0:
            print("%d cases, %d differences, %.0f s" % (len(cases), bad, time.time() - t0), flush=True)
    print("oracle against the reference under random parameter sets: %d cases, %d differences" % (len(cases), bad))
    if n_keep:
        dump("seq_api_fuzz.json", cases[:n_keep])
    return bad


if __name__ == "__main__" and "fuzz" in sys.argv[1:]:
    k = sys.argv.index("fuzz")
    sys.exit(1 if fuzz_golden(int(sys.argv[k + 1]), int(sys.argv[k + 2]) if len(sys.argv) > k + 2 else 0,
                              int(sys.argv[k + 3]) if len(sys.argv) > k + 3 else 20261018,
                              int(sys.argv[k + 4]) if len(sys.argv) > k + 4 else 8,
                              int(sys.argv[k + 5]) if len(sys.argv) > k + 5 else 110) else 0)

if __name__ == "__main__" and "bpp" in sys.argv[1:]:
    bpp_golden()
    sys.exit(0)

if __name__ == "__main__" and "xlong3k" in sys.argv[1:]:
    xlong3k_golden()
    sys.exit(0)

if __name__ == "__main__" and "c3b" in sys.argv[1:]:
    c3b_golden(int(sys.argv[sys.argv.index("c3b") + 1]))
    sys.exit(0)

if __name__ == "__main__" and "c4" in sys.argv[1:]:
    c4_golden(16, verbose=True)
    c4_golden(64)
    c4_golden(256)
    sys.exit(0)

if __name__ == "__main__" and "c3" in sys.argv[1:]:
    c3_golden()
    sys.exit(0)

if __name__ == "__main__" and "xlong" in sys.argv[1:]:
    xlong_golden()
    sys.exit(0)

if __name__ == "__main__" and "long" in sys.argv[1:]:
    long_golden()
    sys.exit(0)

if __name__ == "__main__" and "algos" in sys.argv[1:]:
    algos_golden()
    sys.exit(0)

if __name__ == "__main__" and "cli" not in sys.argv[1:]:
    main()


# --- G5: CLI text of the reference on the bundled example inputs ------------------
CLI_RUNS = {
    "seq_greedynobpp": ["i=inputs/seq_input.fas", "c=greedynobpp", "t=4"],
    "seq_greedynobpp_opts": ["i=inputs/seq_input.fas", "c=greedynobpp", "t=4", "pl=20", "tl=3", "ol=4", "cl=2", "rb=drs", "hr", "rf=26"],
    "seq_fastest_byseq": ["i=inputs/seq_input.fas", "c=fastest", "byseq", "pl=1", "t=4"],
    "seq_alt_msn": ["i=inputs/seq_input.fas", "c=alt", "t=4", "msn=3", "rb=s", "rf=10", "ico"],
    "seq_evalonly": ["i=inputs/seq_input.fas", "c=fastest", "eo", "t=2"],
    "shape_fastest": ["i=inputs/shape_input.fas", "c=fastest", "t=2", "pl=1"],
    "shape_greedynobpp": ["i=inputs/shape_input.fas", "c=greedynobpp", "t=4"],
    "inline_seq": ["s=GGGAAACCCAAAGGGUUUCCC", "c=greedynobpp", "t=2"],
    "ali_default": ["i=inputs/ali_input.afa", "a", "t=4"],
    "ali_verbose": ["i=inputs/ali_input.afa", "a", "v", "t=4"],
    "ali_step3_i": ["i=inputs/ali_input.afa", "a", "s3=i", "fl=0.5", "ll=1", "t=4"],
    "ali_step3_1": ["i=inputs/ali_input.afa", "a", "s3=1", "t=4"],
    "ali_demo": ["i=inputs/demo.afa", "a", "t=2"],
    "ali_as_single_fastest": ["i=inputs/ali_input.afa", "c=fastest", "pl=1", "byseq", "t=4", "if=q"],
    "seq_nobpp": ["i=inputs/seq_input.fas", "c=nobpp", "t=4"],                     # G + Nussinov + Edmonds + Hungarian sets
    "shape_nobpp_opts": ["i=inputs/shape_input.fas", "c=nobpp", "t=4", "pl=50", "tl=4", "ol=6", "cl=2", "rb=dr", "ll=1"],
    "seq_edmondsnobpp": ["i=inputs/seq_input.fas", "c=edmondsnobpp", "t=2"],
    "ali_nobpp": ["i=inputs/ali_input.afa", "a", "c=nobpp", "t=4"],                # alignment mode, step 2 with all five sets
    "seq_greedynobpp_ent": ["i=inputs/seq_input.fas", "c=greedynobpp", "ent", "t=4"],          # the entropy line
    "seq_fastest_byseq_ent": ["i=inputs/shape_input.fas", "c=fastest", "byseq", "pl=1", "ent", "t=2"],
    "ali_demo_ent_verbose": ["i=inputs/demo.afa", "a", "v", "ent", "t=2"],
    # the CLI's DEFAULT configs (def.conf: 12 sets, 7 of them with bpp) -- with tests/fake_rna.py as the RNA module
    "seq_default_fakerna": ["i=inputs/seq_input.fas", "t=4"],
    "inline_default_fakerna": ["s=GGGAAACCCAAAGGGUUUCCCAAAGGCGAAAGCC", "t=2"],
    "shape_greedy_fakerna": ["i=inputs/shape_input.fas", "c=greedy", "t=2", "pl=20"],
    # BASELINE config 4 shape: synthetic alignments from workloads.config4 (inputs written by `make_golden.py c4`)
    # error behaviour: the 12th of 13 entries has a malformed reactivities line.  Streaming mode has printed the 11 before
    # it; byseq prints in groups of threads * 10 entries (SQUARNA.py:887-935): the first 10 with t=1, none with t=2
    "bad_entry_error": ["i=inputs/bad_entry.fas", "c=greedynobpp", "pl=5", "t=2"],
    "bad_entry_byseq_t1_error": ["i=inputs/bad_entry.fas", "c=greedynobpp", "pl=5", "byseq", "t=1"],
    "bad_entry_byseq_t2_error": ["i=inputs/bad_entry.fas", "c=fastest", "pl=1", "byseq", "t=2"],
    "ali_c4_64x400": ["i=inputs/ali_c4_64x400.afa", "a", "t=3"],
    "ali_c4_256x400": ["i=inputs/ali_c4_256x400.afa", "a", "t=3"],
}


def cli_golden():
    import subprocess
    man = {}
    only = set(sys.argv[sys.argv.index("cli") + 1:])          # `cli name ...`: regenerate just these texts
    for name, argv in CLI_RUNS.items():
        env = dict(os.environ)
        if name.endswith("_fakerna"):
            env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "_fake_site"), os.path.dirname(os.path.dirname(HERE))])
        if only and name not in only:
            man[name] = argv
            continue
        out = subprocess.run([sys.executable, os.path.join(REF, "SQUARNA.py")] + argv, cwd=HERE,
                             capture_output=True, text=True, env=env)
        if name.endswith("_error"):               # a malformed entry: what was printed before the exception, and its type
            assert out.returncode != 0, name
            with open(os.path.join(HERE, "cli", name + ".err"), "w") as f:
                import re
                f.write([m.group(1) for m in (re.match(r"(\w+(?:Error|Exception))\b", ln) for ln in out.stderr.split("\n")) if m][-1] + "\n")
        else:
            assert out.returncode == 0, (name, out.stderr[-2000:])
        with open(os.path.join(HERE, "cli", name + ".txt"), "w") as f:
            f.write(out.stdout)
        man[name] = argv
        print(name, len(out.stdout), "chars")
    dump("cli_manifest.json", man)


if __name__ == "__main__" and "cli" in sys.argv[1:]:
    cli_golden()

"""Randomised comparison of the whole host side with the REAL reference (needs /root/reference, so it runs in the build
container only; not collected by pytest): random input files and random command-line options through the reference's
own `SQUARNA.Main` and through `squarna_b200.SQUARNA.Main` with the oracle standing in for the GPU calls (the stand-ins
of tests/test_host_python.py).  Single-sequence mode (default-format and FASTA inputs with reactivity / restraint /
reference lines, separators, gaps) and alignment mode (small random alignments with default lines).

    python -m tests.fuzz_cli [seed] [cases]
"""
import contextlib
import io
import os
import random
import sys
import tempfile

REF = "/root/reference/src/SQUARNA"


def load_reference():
    sys.path.insert(0, REF)                  # (top-level module names, as tests/golden/make_golden.py imports them: the
    import SQUARNA as mod                    #  reference's worker functions must stay picklable for its Pool)
    return mod


def rand_seq(rng, n, alphabet="ACGU"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def rand_restraints(rng, seq):
    out = ["."] * len(seq)
    for k in range(len(seq)):
        x = rng.random()
        if seq[k] in ";&-.~":
            continue
        out[k] = "_" if x < 0.06 else "/" if x < 0.08 else "\\" if x < 0.10 else "+" if x < 0.11 else "."
    if rng.random() < 0.5 and len(seq) > 20:              # a planted stem as brackets
        i = rng.randrange(0, len(seq) - 14)
        j = rng.randrange(i + 9, len(seq))
        ln = rng.randint(1, 3)
        cells = list(range(i, i + ln)) + list(range(j - ln + 1, j + 1))
        if j - i >= 2 * ln + 3 and all(seq[c] not in ";&-.~" for c in cells):
            for q in range(ln):
                out[i + q], out[j - q] = "(", ")"
    return "".join(out)


def rand_reacts(rng, seq, fmt):
    n = len(seq)
    if fmt == 26:
        return "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") if rng.random() > 0.05 else "?" for _ in range(n))
    if fmt == 10:
        return "".join(rng.choice("0123456789") if rng.random() > 0.05 else "?" for _ in range(n))
    if fmt == 3:
        return "".join(rng.choice("_+#") if rng.random() > 0.05 else "?" for _ in range(n))
    return " ".join("%.2f" % rng.random() if rng.random() > 0.1 else "-999" for _ in range(n))


def rand_single_input(rng, fmt):
    """default-format input: name, sequence, then optional reactivities / restraints / reference lines"""
    lines = []
    for k in range(rng.randint(1, 5)):
        n = rng.randint(6, 90)
        seq = rand_seq(rng, n, rng.choice(["ACGU", "ACGU", "ACGUT", "ACGUacgu", "ACGU-", "ACGUN"]))
        if seq[0] == "-":
            seq = "A" + seq[1:]
        if rng.random() < 0.2 and n > 12:
            p = rng.randrange(3, n - 3)
            seq = seq[:p] + rng.choice(";&") + seq[p + 1:]
        lines.append(">case%d %s" % (k, rng.choice(["", "with a description"])))
        lines.append(seq)
        # positional lines after the sequence: reactivities, restraints, reference (an empty line skips one)
        extra = [rand_reacts(rng, seq, fmt) if rng.random() < 0.5 else "", rand_restraints(rng, seq) if rng.random() < 0.5 else ""]
        ref = ""
        if rng.random() < 0.3:
            ref = ["."] * n
            i, j = sorted(rng.sample(range(n), 2))
            if j - i > 4 and seq[i] not in ";&-" and seq[j] not in ";&-":
                ref[i], ref[j] = "(", ")"
            ref = "".join(ref)
        extra.append(ref)
        while extra and not extra[-1]:
            extra.pop()
        if rng.random() < 0.04 and extra:                      # a malformed line now and then: the error paths
            extra[rng.randrange(len(extra))] = "(((" if rng.random() < 0.5 else "xyz" * 3
        lines += extra
    return "\n".join(lines) + "\n"


def rand_alignment(rng):
    n_cols, anc_len = rng.randint(40, 90), 0
    anc_len = rng.randint(n_cols * 2 // 3, n_cols)
    anc = list(rand_seq(rng, anc_len))
    comp = {"A": "U", "U": "A", "G": "C", "C": "G"}
    pos, partner = 2, {}
    while pos + 22 < anc_len:
        ln = rng.randint(4, 7)
        i, j = pos, pos + 2 * ln + 4
        for q in range(ln):
            anc[j - q] = comp[anc[i + q]]
            partner[i + q], partner[j - q] = j - q, i + q
        pos = j + rng.randint(3, 8)
    cols = sorted(rng.sample(range(n_cols), anc_len))
    rows = []
    for _ in range(rng.randint(3, 9)):
        seq = anc[:]
        for p in range(anc_len):
            if rng.random() < 0.12:
                seq[p] = rng.choice("ACGU")
                if p in partner:
                    seq[partner[p]] = comp[seq[p]]
        row = ["-"] * n_cols
        for p in range(anc_len):
            if rng.random() > 0.04:
                row[cols[p]] = seq[p]
        rows.append("".join(row))
    lines = []
    if rng.random() < 0.5:                                   # default lines: reactivities, restraints, reference
        ref = ["."] * n_cols
        for p, q in partner.items():
            if p < q:
                ref[cols[p]], ref[cols[q]] = "(", ")"
        lines += ["?" * n_cols, "." * n_cols, "".join(ref)]
    for k, r in enumerate(rows):
        lines += [">row%d" % k, r]
    return "\n".join(lines) + "\n"


BPP_CONFIGS = False          # campaign(..., bpp=True): also the configs whose sets ask ViennaRNA for pair probabilities


def rand_args(rng, alignment):
    if alignment:
        args = ["a"]
        if rng.random() < 0.4:
            args.append("v")
        if rng.random() < 0.5:
            args.append("s3=" + rng.choice(["i", "u", "1", "2"]))
        if rng.random() < 0.3:
            args.append("fl=%.2f" % rng.choice([0.0, 0.3, 0.5, 0.8]))
        if rng.random() < 0.3:
            args.append("ll=%d" % rng.choice([1, 2, 3]))
        if rng.random() < 0.3:
            args.append("c=" + rng.choice(["ali", "fastest", "greedynobpp"]))
        if rng.random() < 0.3:
            args.append("pl=%d" % rng.choice([1, 5, 50]))
        if rng.random() < 0.15:
            args.append("ent")
        if rng.random() < 0.15:
            args.append("ico")
        return args, None
    fmt = rng.choice([None, 3, 10, 26])
    confs = ["fastest", "greedynobpp", "alt", "500nobpp", "1000nobpp", "nobpp", "edmondsnobpp", "hungariannobpp", "nussinovnobpp"]
    if BPP_CONFIGS:
        confs = ["def", "greedy", "500", "1000", "edmonds", "hungarian", "nussinov", "greedynobpp"]
    args = ["c=" + rng.choice(confs)]
    if fmt:
        args.append("rf=%d" % fmt)
    for flag, p in (("byseq", 0.3), ("hr", 0.2), ("ico", 0.15), ("eo", 0.08), ("iw", 0.2), ("ent", 0.12)):
        if rng.random() < p:
            args.append(flag)
    for key, choices, p in (("pl", [1, 3, 20, 100], 0.6), ("tl", [1, 3, 5], 0.3), ("ol", [1, 2, 4], 0.3), ("cl", [1, 2, 3], 0.3),
                            ("rb", ["r", "s", "d", "rs", "dsr", "sd"], 0.4), ("msn", [1, 3], 0.15), ("ll", [1, 2], 0.15),
                            ("algo", ["G", "N", "GE", "H"], 0.12), ("pr", ["defG1", "defG2,defG1", "fastestG", "nosuchset"], 0.1),
                            ("if", ["qtrf", "qrtf", "qtr", "qr", "qf", "qt"], 0.15)):
        if rng.random() < p:
            args.append("%s=%s" % (key, rng.choice(choices)))
    return args, fmt


def run_main(main, args, cwd):
    buf, err = io.StringIO(), io.StringIO()
    here = os.getcwd()
    os.chdir(cwd)
    outcome = "ok"
    try:
        with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(err):
            main(args)
    except SystemExit as e:
        outcome = "exit %r" % (e.code,)
    except Exception as e:                     # noqa: BLE001 -- the kind of failure is part of the comparison
        outcome = "raise " + type(e).__name__
    finally:
        os.chdir(here)
    return outcome, buf.getvalue()


def campaign(seed, cases, verbose=True, bpp=False):
    global BPP_CONFIGS
    BPP_CONFIGS = bpp
    from _pytest.monkeypatch import MonkeyPatch
    from squarna_b200 import SQRNdbnali as A
    from squarna_b200 import SQRNdbnseq as S
    from squarna_b200 import SQUARNA as CLI
    from tests import test_host_python as H
    ref = load_reference()
    mp = MonkeyPatch()
    H._OracleContext.fast_predict = H._oracle_fast_predict
    H._stand_in(mp)
    mp.setattr(A, "_yield_many", H._oracle_yield_many)
    if bpp:                                   # tests/fake_rna.py stands in for ViennaRNA on both sides
        from tests import fake_rna
        sys.modules["RNA"] = fake_rna
        S.set_rna_module(fake_rna)
    rng = random.Random(seed)
    bad, outcomes, n_chars = 0, {}, 0
    try:
        with tempfile.TemporaryDirectory() as tmp:
            for k in range(cases):
                alignment = rng.random() < 0.35
                args, fmt = rand_args(rng, alignment)
                text = rand_alignment(rng) if alignment else rand_single_input(rng, fmt)
                with open(os.path.join(tmp, "in.fas"), "w") as f:
                    f.write(text)
                args = ["i=in.fas"] + args
                def ref_main(argv):                       # the reference reads sys.argv (SQUARNA.py:1012)
                    saved = sys.argv
                    sys.argv = ["SQUARNA"] + argv
                    try:
                        ref.Main()
                    finally:
                        sys.argv = saved
                want = run_main(ref_main, args + ["t=1"], tmp)
                got = run_main(CLI.Main, args + ["t=1"], tmp)
                outcomes[want[0].split()[0]] = outcomes.get(want[0].split()[0], 0) + 1
                n_chars += len(want[1])
                if got != want:
                    bad += 1
                    print("DIFFERENCE #%d: args %r\n--- input\n%s--- reference (%s)\n%s--- ours (%s)\n%s" %
                          (bad, args, text, want[0], want[1][:3000], got[0], got[1][:3000]), flush=True)
                    if bad >= 3:
                        break
                elif verbose and k % 20 == 19:
                    print("%d cases, %d differences" % (k + 1, bad), flush=True)
    finally:
        mp.undo()
        S.set_rna_module(None)
    print("CLI against the reference: %d differences; reference outcomes %r, %d characters of output compared" % (bad, outcomes, n_chars))
    return bad


if __name__ == "__main__":
    sys.exit(1 if campaign(int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 100,
                           bpp="bpp" in sys.argv[3:]) else 0)

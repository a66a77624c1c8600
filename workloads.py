"""Seeded synthetic inputs of BASELINE.json's configs 2-5, shared by bench.py, the
tests and tests/golden/make_golden.py.  Bases are i.i.d. uniform over ACGU; everything is a pure
function of (size, seed), so the GPU box regenerates exactly what the build container saw.

    config2  1 M sequences U{60..200}                         byseq pl=1 c=fastest.conf
    config3  100 k sequences U{300..1500}, rf=26 reactivity letters (3 % '?'), restraints
             (5 % '_', 1 % '/', 1 % '\\', 0-2 planted canonical stems of 3-6 bp as brackets),
             G sets by length (greedynobpp < 500, 500nobpp G sets 500-999, 1000nobpp >= 1000), pl=100
    config4  alignment 2000 x 400 shaped like examples/ali_input.afa (3 default lines), `a`, ali.conf
    config5  10 k sequences U{2900..5000}, 1000nobpp.conf G set, pl=1
"""
import numpy as np

SEED = 20261017
_ACGU = np.frombuffer(b"ACGU", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in ("AU", "UA", "GC", "CG"):
    _COMP[ord(_a)] = ord(_b)


def _csr(rng, n, lo, hi):
    lens = rng.integers(lo, hi + 1, size=n, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    sym = _ACGU[rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)]
    return np.ascontiguousarray(sym), off, lens


def config2(n, seed=SEED):
    """(symbols u8, offsets i64, lengths): n plain sequences of 60..200 nt"""
    return _csr(np.random.default_rng(seed), n, 60, 200)


def config5(n, seed=SEED, lo=2900, hi=5000):
    """(symbols u8, offsets i64, lengths): n plain sequences of 2900..5000 nt"""
    return _csr(np.random.default_rng(seed + 5), n, lo, hi)


def config3_conf(length):
    """the reference's autoconfig by length (SQUARNA.py:870-878) restricted to the bpp-free G sets"""
    return "greedynobpp" if length < 500 else "500nobpp" if length < 1000 else "1000nobpp"


def config3(n, seed=SEED, lo=300, hi=1500):
    """list of (sequence, reactivity letters, restraint string) -- all str"""
    rng = np.random.default_rng(seed + 3)
    sym, off, lens = _csr(rng, n, lo, hi)
    sym = sym.copy()
    total = int(off[-1])
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)[rng.integers(0, 26, size=total)]
    letters = np.where(rng.random(total) < 0.03, np.uint8(ord("?")), letters)
    x = rng.random(total)
    rest = np.full(total, ord("."), dtype=np.uint8)
    rest[x < 0.07] = ord("\\")
    rest[x < 0.06] = ord("/")
    rest[x < 0.05] = ord("_")
    nplant = rng.integers(0, 3, size=n)
    out = []
    for b in range(n):
        o, N = int(off[b]), int(lens[b])
        s, r = sym[o:o + N], rest[o:o + N]
        for _ in range(int(nplant[b])):                       # a planted canonical stem, given as a bracket restraint
            ln = int(rng.integers(3, 7))
            i = int(rng.integers(5, N // 2 - 10))
            j = int(rng.integers(N // 2 + 10, N - 5))
            if (r[i:i + ln] == ord("(")).any() or (r[j - ln + 1:j + 1] == ord(")")).any():
                continue
            s[j - ln + 1:j + 1] = _COMP[s[i:i + ln]][::-1]
            r[i:i + ln] = ord("(")
            r[j - ln + 1:j + 1] = ord(")")
        out.append((s.tobytes().decode(), letters[o:o + N].tobytes().decode(), r.tobytes().decode()))
    return out


def config4(n_seqs, anc_len=300, n_cols=400, seed=SEED):
    """(rows, reference dbn): an ancestor with planted hairpins, descendants with 10 % substitutions
    (compensatory inside stems), 2 % deletions and shared gap columns up to n_cols"""
    rng = np.random.default_rng(seed + 4)
    anc = _ACGU[rng.integers(0, 4, size=anc_len)].copy()
    partner = np.full(anc_len, -1, dtype=np.int64)
    pos = 5
    while pos + 40 < anc_len:                                   # hairpins of 6-9 bp around 5-nt loops
        ln = int(rng.integers(6, 10))
        i, j = pos, pos + 2 * ln + 4
        anc[j - ln + 1:j + 1] = _COMP[anc[i:i + ln]][::-1]
        partner[i:i + ln] = np.arange(j, j - ln, -1)
        partner[j - ln + 1:j + 1] = np.arange(i + ln - 1, i - 1, -1)
        pos = j + int(rng.integers(4, 11))
    cols = np.sort(rng.choice(n_cols, size=anc_len, replace=False))
    ref = np.full(n_cols, ord("."), dtype=np.uint8)
    five = np.flatnonzero((partner >= 0) & (np.arange(anc_len) < partner))
    ref[cols[five]] = ord("(")
    ref[cols[partner[five]]] = ord(")")
    rows = []
    for _ in range(n_seqs):
        seq = anc.copy()
        for p in np.flatnonzero(rng.random(anc_len) < 0.10):
            seq[p] = _ACGU[rng.integers(0, 4)]
            if partner[p] >= 0:
                seq[partner[p]] = _COMP[seq[p]]
        row = np.full(n_cols, ord("-"), dtype=np.uint8)
        keep = rng.random(anc_len) > 0.02
        row[cols[keep]] = seq[keep]
        rows.append(row.tobytes().decode())
    return rows, ref.tobytes().decode()


def config4_text(rows, ref):
    """the alignment as an input file in the reference's default format: the three default lines
    (reactivities '?', restraints '.', reference) before the first '>' (SQUARNA.py:184-191)"""
    n_cols = len(ref)
    lines = ["?" * n_cols, "." * n_cols, ref]
    for k, r in enumerate(rows):
        lines.append(">seq%d" % k)
        lines.append(r)
    return "\n".join(lines) + "\n"


def algorithmic_bytes(lens, n_structs=1, reacts=False, restraints=False):
    """ceil(N/4) + 8 + N [reacts] + N [restraints] + S (N + 32) + N per sequence"""
    lens = np.asarray(lens, dtype=np.int64)
    per = (lens + 3) // 4 + 8 + lens * (int(reacts) + int(restraints)) + n_structs * (lens + 32) + lens
    return int(per.sum())

#!/usr/bin/env python
"""Throughput of the BASELINE.json configs other than the headline one, at reduced size
(the headline config 2 is bench.py).  One JSON line per leg; run on the GPU box:

    python scripts/bench_configs.py c5 --n 16        # rRNA-scale, 1000nobpp G set, pl=1
    python scripts/bench_configs.py c3 --n 64        # 300..1500 nt, reactivities + restraints, G sets by length, pl=100
    python scripts/bench_configs.py mid --n 2000     # 300..1500 nt plain, fastest.conf pl=1 (CTA teams)
"""
import argparse, json, os, random, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from squarna_b200 import SQRNdbnseq as S, SQUARNA as CLI, _lib           # noqa: E402
from squarna_b200._abi import pack_sequences                              # noqa: E402

CONF = os.path.join(ROOT, "squarna_b200")


def gsets(name):
    names, ps = CLI.ParseConfig(os.path.join(CONF, name + ".conf"))
    return [p for p in ps if p["algorithms"] == {"G"} and not p.get("bpp", 0)]


def rand_seqs(rng, n, lo, hi):
    return ["".join(rng.choice("ACGU") for _ in range(rng.randint(lo, hi))) for _ in range(n)]


def leg_tail(name, seqs, ps, oracle_n):
    ctx = S.get_context(0)
    if os.environ.get("SQRN_CLUSTER"):            # 1: never a cluster, 2/4/8/16: always that size (default: automatic)
        ctx.set_cluster(int(os.environ["SQRN_CLUSTER"]))
        name += " cluster=" + os.environ["SQRN_CLUSTER"]
    sym, off = pack_sequences(seqs)
    k = int(np.argmax(np.diff(off)))                            # warm-up on the longest sequence: module load, parameter
    ctx.fast_predict(ps, sym[int(off[k]):int(off[k + 1])], np.array([0, off[k + 1] - off[k]], dtype=np.int64))   # digest, scratch
    t0 = time.perf_counter()
    dbn, sc, nst = ctx.fast_predict(ps, sym, off)
    dt = time.perf_counter() - t0
    st = ctx.stats()
    lens = np.diff(off).astype(np.float64)
    line = {"leg": name, "n_seqs": len(seqs), "len_min": int(lens.min()), "len_max": int(lens.max()), "seconds": dt,
            "seq_per_s": len(seqs) / dt, "nt2_per_s": float((lens ** 2).sum()) / dt, "kernel_ms": st["kernel_ms"],
            "optimal_calls": st["optimal_calls"], "stems_mean": float(nst.mean())}
    if oracle_n:
        from oracle import oracle as O
        k = min(oracle_n, len(seqs))
        t0 = time.perf_counter()
        odbn, osc, onst = O.predict_batch_simple(sym[:int(off[k])], off[:k + 1], [ps], poollim=1, nthreads=os.cpu_count() or 1)
        line["oracle_seconds_for_%d" % k] = time.perf_counter() - t0
        line["parity"] = bool(np.array_equal(onst, nst[:k]) and np.array_equal(osc, sc[:k]))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("leg")
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--oracle", type=int, default=0, help="check the first K sequences against the CPU oracle")
    ap.add_argument("--lo", type=int, default=0)
    ap.add_argument("--hi", type=int, default=0)
    a = ap.parse_args()
    rng = random.Random(20261017)
    if a.leg == "c5":
        leg_tail("c5 1000nobpp G pl=1", rand_seqs(rng, a.n, a.lo or 2900, a.hi or 5000), gsets("1000nobpp")[0], a.oracle)
    elif a.leg == "mid":
        leg_tail("mid fastest pl=1", rand_seqs(rng, a.n, a.lo or 300, a.hi or 1500), gsets("fastest")[0], a.oracle)
    elif a.leg == "c4":
        leg_alignment(a.n, a.lo or 300, a.hi or 400)
    elif a.leg == "c3":
        seqs = rand_seqs(rng, a.n, a.lo or 300, a.hi or 1500)
        entries = []
        for s in seqs:
            n = len(s)
            reacts = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") if rng.random() > 0.03 else "?" for _ in range(n))
            rest = ["."] * n
            for k in range(n):
                x = rng.random()
                rest[k] = "_" if x < 0.05 else "/" if x < 0.06 else "\\" if x < 0.07 else "."
            entries.append((s, reacts, "".join(rest), None))
        groups = {"greedynobpp": [], "500nobpp": [], "1000nobpp": []}
        for e in entries:
            groups["greedynobpp" if len(e[0]) < 500 else "500nobpp" if len(e[0]) < 1000 else "1000nobpp"].append(e)
        S.predict_many(entries[:1], gsets("fastest"), poollim=1)
        for conf, es in groups.items():
            if not es:
                continue
            t0 = time.perf_counter()
            out = S.predict_many(es, gsets(conf), poollim=100)
            dt = time.perf_counter() - t0
            st = S.get_context(0).stats()
            lens = np.array([len(e[0]) for e in es], dtype=np.float64)
            print(json.dumps({"leg": "c3 " + conf + " pl=100", "n_seqs": len(es), "seconds": dt, "seq_per_s": len(es) / dt,
                              "nt2_per_s": float((lens ** 2).sum()) / dt, "launches": st["launches"],
                              "optimal_calls": st["optimal_calls"], "kernel_ms": st["kernel_ms"],
                              "structs_mean": float(np.mean([len(o[1]) for o in out]))}), flush=True)


def make_alignment(rng, n_seqs, anc_len, n_cols):
    """ancestor with a planted nested structure; descendants with 10 % substitutions (compensatory in
    stems) and gap columns up to n_cols"""
    comp = {"A": "U", "U": "A", "G": "C", "C": "G"}
    anc = [rng.choice("ACGU") for _ in range(anc_len)]
    pairs = []
    pos = 5
    while pos + 40 < anc_len:                       # hairpins of 6-9 bp with 5-nt loops
        ln = rng.randint(6, 9)
        i, j = pos, pos + 2 * ln + 4
        for k in range(ln):
            anc[j - k] = comp[anc[i + k]]
            pairs.append((i + k, j - k))
        pos = j + rng.randint(4, 10)
    gapcols = sorted(rng.sample(range(n_cols), n_cols - anc_len))
    colmap = [c for c in range(n_cols) if c not in set(gapcols)]
    ref = ["."] * n_cols
    for v, w in pairs:
        ref[colmap[v]], ref[colmap[w]] = "(", ")"
    rows = []
    partner = dict(pairs); partner.update({w: v for v, w in pairs})
    for _ in range(n_seqs):
        seq = list(anc)
        for p in range(anc_len):
            if rng.random() < 0.10:
                seq[p] = rng.choice("ACGU")
                if p in partner:
                    seq[partner[p]] = comp[seq[p]]
        row = ["-"] * n_cols
        for p, c in enumerate(colmap):
            row[c] = seq[p] if rng.random() > 0.02 else "-"
        rows.append("".join(row))
    return rows, "".join(ref)


def leg_alignment(n_seqs, anc_len, n_cols):
    import io, tempfile
    rng = random.Random(20261017)
    rows, ref = make_alignment(rng, n_seqs, anc_len, n_cols)
    with tempfile.NamedTemporaryFile("w", suffix=".afa", delete=False) as f:
        f.write("?" * n_cols + "\n" + "." * n_cols + "\n" + ref + "\n")
        for k, r in enumerate(rows):
            f.write(">seq%d\n%s\n" % (k, r))
        path = f.name
    S.get_context(0)
    sink = io.StringIO()
    t0 = time.perf_counter()
    CLI.Predict(inputfile=path, alignment=True, write_to=sink)
    dt = time.perf_counter() - t0
    text = sink.getvalue().split("\n")
    tail = [ln for ln in text if ln.strip()][-6:]
    print(json.dumps({"leg": "c4 alignment ali.conf", "n_seqs": n_seqs, "columns": n_cols, "seconds": dt,
                      "seq_per_s": n_seqs / dt, "output_tail": tail}), flush=True)


if __name__ == "__main__":
    main()

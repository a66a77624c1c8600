"""Host-side sharding logic: the plan, the CSR gather/scatter, and the N > 1 path with
two `gloo` ranks on CPU.  The per-rank "predictor" is a stub (a checksum per sequence and a
per-position transform): what is tested here is that sharded results come back in input order."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from squarna_b200 import sharding as SH


def _batch(seed, n, lo, hi):
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    sym = np.frombuffer(b"ACGU", dtype=np.uint8)[rng.integers(0, 4, size=int(off[-1]))]
    return np.ascontiguousarray(sym), off


def _stub_predict(sym, off):
    """stands in for the GPU: one checksum per sequence, one transformed byte per position"""
    n = len(off) - 1
    score = np.array([int(sym[off[b]:off[b + 1]].astype(np.int64).sum()) * 31 + (off[b + 1] - off[b]) for b in range(n)])
    return score.reshape(n, 1), (sym ^ 0x20).astype(np.uint8)


def test_plan_covers_everything_once_and_balances():
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 8):
        lens = rng.integers(60, 201, size=5000)
        plan = SH.shard_plan(lens, world)
        allidx = np.concatenate(plan)
        assert sorted(allidx.tolist()) == list(range(len(lens)))
        assert max(len(p) for p in plan) - min(len(p) for p in plan) <= 1
        assert SH.plan_imbalance(lens, plan, 2.0) < 1.01
        for p in plan:                                      # longest first inside a queue
            assert (np.diff(lens[p]) <= 0).all()
    # ragged: fewer sequences than queues, empty batch
    assert [len(p) for p in SH.shard_plan([5, 7], 4)] == [1, 1, 0, 0]
    assert [len(p) for p in SH.shard_plan([], 2)] == [0, 0]
    heavy = np.array([5000] + [100] * 999)
    assert SH.plan_imbalance(heavy, SH.shard_plan(heavy, 8), 3.0) > 1.0   # one rRNA dominates: reported, not hidden


def test_contiguous_plan_balances_cost():
    rng = np.random.default_rng(4)
    lens = rng.integers(60, 201, size=100000)
    for world in (1, 2, 4, 8):
        cuts = SH.contiguous_plan(lens, world)
        assert cuts[0] == 0 and cuts[-1] == len(lens) and (np.diff(cuts) >= 0).all() and len(cuts) == world + 1
        cost = np.array([(lens[a:b].astype(np.float64) ** 2).sum() for a, b in zip(cuts[:-1], cuts[1:])])
        assert cost.max() / cost.mean() < 1.001
    assert SH.contiguous_plan([], 4).tolist() == [0, 0, 0, 0, 0]
    assert SH.contiguous_plan([10], 3).tolist()[-1] == 1


def test_run_sharded_keeps_input_order():
    """the one-host-thread-per-GPU dealer of predict_many / YieldStems: every item once, results in input order,
    errors of a worker surface in the caller"""
    from squarna_b200 import SQRNdbnseq as S
    items = ["x" * n for n in np.random.default_rng(5).integers(1, 400, size=257).tolist()]
    seen = []

    def fn(sub, dev):
        seen.append((dev, len(sub)))
        return [(dev, len(x)) for x in sub]

    out = S.run_sharded(fn, items, [len(x) for x in items], [0, 1, 2])
    assert [o[1] for o in out] == [len(x) for x in items] and sorted(d for d, _ in seen) == [0, 1, 2]
    assert sum(k for _, k in seen) == len(items)

    def bad(sub, dev):
        raise ValueError("boom")
    with pytest.raises(ValueError, match="boom"):
        S.run_sharded(bad, items, [len(x) for x in items], [0, 1])


def test_csr_take_put_round_trip():
    sym, off = _batch(2, 300, 0, 90)                        # includes empty sequences
    idx = np.random.default_rng(3).permutation(300)[:170]
    sub, soff = SH.take_csr(sym, off, idx)
    for k, b in enumerate(idx):
        assert np.array_equal(sub[soff[k]:soff[k + 1]], sym[off[b]:off[b + 1]])
    back = np.zeros_like(sym)
    SH.put_csr(back, off, idx, sub, soff)
    rest = np.setdiff1d(np.arange(300), idx)
    sub2, soff2 = SH.take_csr(sym, off, rest)
    SH.put_csr(back, off, rest, sub2, soff2)
    assert np.array_equal(back, sym)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, seed, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sym, off = _batch(seed, 1200, 1, 180)
        idx = SH.rank_shard(np.diff(off), rank, world)
        sub, soff = SH.take_csr(sym, off, idx)
        score, per_pos = _stub_predict(sub, soff)
        seq_full, pos_full = SH.gather_to_root(idx, {"score": score}, {"dbn": per_pos}, soff, off, rank, world, dist)
        if rank == 0:
            np.savez(out_path, score=seq_full["score"], dbn=pos_full["dbn"])
        else:
            assert seq_full is None and pos_full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_match_single_process(tmp_path):
    out = str(tmp_path / "merged.npz")
    mp.spawn(_rank_main, args=(2, _free_port(), 7, out), nprocs=2, join=True)
    sym, off = _batch(7, 1200, 1, 180)
    score, per_pos = _stub_predict(sym, off)
    got = np.load(out)
    assert np.array_equal(got["score"], score)
    assert np.array_equal(got["dbn"], per_pos)


@pytest.mark.gpu
def test_multigpu_equals_single_context(gpu_ctx):
    from tests import common as T
    from squarna_b200._abi import pack_sequences
    seqs = T.rand_seqs(41, 3000, 1, 200)
    sym, off = pack_sequences(seqs)
    want = gpu_ctx.fast_predict(T.FASTEST, sym, off)
    m = SH.MultiGPU()
    try:
        got = m.fast_predict(T.FASTEST, sym, off)
    finally:
        m.close()
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # the dealt (non-contiguous) path: a few sequences above 320 nt send the batch through shard_plan
    seqs2 = T.rand_seqs(42, 400, 1, 200) + T.rand_seqs(43, 6, 330, 700)
    sym2, off2 = pack_sequences(seqs2)
    want2 = gpu_ctx.fast_predict(T.FASTEST, sym2, off2)
    m = SH.MultiGPU()
    try:
        got2 = m.fast_predict(T.FASTEST, sym2, off2)
    finally:
        m.close()
    for a, b in zip(got2, want2):
        assert np.array_equal(a, b)


@pytest.mark.gpu
def test_predict_many_over_all_gpus_equals_one(gpu_ctx):
    """predict_many(devices=None) deals the entries to every visible GPU (one host thread each); same predictions,
    same order as on one device -- pools, restraints, reactivities included"""
    import random
    from tests import common as T
    from squarna_b200 import SQRNdbnseq as S
    rng = random.Random(77)
    cases = [T.rand_case(rng, 20, 160) for _ in range(96)]
    entries = [(c[0], c[1], c[2], None) for c in cases]
    one = S.predict_many(entries, [T.DEFG1, T.DEFG2], poollim=20, device=0)
    every = S.predict_many(entries, [T.DEFG1, T.DEFG2], poollim=20, devices=None)
    assert len(one) == len(every) == len(entries)
    for a, b in zip(one, every):
        assert T.same_prediction((a[0], a[1]), (b[0], b[1]))

"""Sharding of a batch of independent sequences over the GPUs of one box.

The reference parallelises with ``multiprocessing.Pool(threads).imap`` over sequences
(SQUARNA.py:887-935 ``byseq``; SQRNdbnali.py:222-237, 382-392) and relies on the ordered
``imap`` for output order.  Sequences are independent, so here they are dealt to per-GPU work
queues by length (the cost of a sequence grows like N^2 .. N^3) and the results are written
back by original index: no collective is needed on the data path.

Two drivers share the same plan:
  * ``MultiGPU``     one process, one host thread per GPU (ctypes drops the GIL during calls);
  * ``rank_shard`` / ``gather_to_root``  one process per GPU under ``torch.distributed``
    (``torchrun``), results gathered on rank 0 -- the layout bench.py and the -m "not gpu"
    gloo test use.
"""
import threading

import numpy as np


def shard_plan(lengths, world, exponent=2.0):
    """Deal sequences to `world` queues so that the queues carry nearly the same sum of N**exponent
    (the cost of a sequence: ~N^2 for short ones, ~N^3 above a few hundred nt): longest first, each
    to the queue with the smallest load so far (LPT).  Sequences of equal length are dealt in blocks
    -- with a million short sequences the loop runs over the distinct lengths, not over the
    sequences -- and the queues end up with the same count (+-1) per length.
    Returns a list of int64 index arrays; inside a queue the order is longest first, which is also
    the order the persistent kernels want (long items first, short ones fill the tail)."""
    lengths = np.asarray(lengths).astype(np.int64)
    n = len(lengths)
    order = np.argsort(-lengths, kind="stable")
    if world <= 1 or n == 0:
        return [order] + [np.zeros(0, dtype=np.int64) for _ in range(max(world, 1) - 1)]
    sl = lengths[order]
    starts = np.flatnonzero(np.r_[True, sl[1:] != sl[:-1]])           # one block per distinct length
    ends = np.r_[starts[1:], n]
    load = np.zeros(world)
    queue = np.empty(n, dtype=np.int64)
    for a, b in zip(starts.tolist(), ends.tolist()):
        cost = float(sl[a]) ** exponent
        by_load = np.argsort(load, kind="stable")                     # lightest queue first
        cnt = b - a
        share = np.full(world, cnt // world, dtype=np.int64)
        share[:cnt % world] += 1                                      # the extra ones go to the lightest queues
        queue[a:b] = np.repeat(by_load, share)
        load[by_load] += share * cost
    return [order[queue == q] for q in range(world)]


def plan_imbalance(lengths, plan, exponent=2.0):
    """max over queues of sum(N**exponent) divided by the mean: 1.0 = perfectly balanced"""
    cost = np.asarray(lengths, dtype=np.float64) ** exponent
    loads = np.array([cost[idx].sum() for idx in plan])
    return float(loads.max() / loads.mean()) if loads.mean() > 0 else 1.0


def take_csr(values, offsets, idx):
    """rows `idx` of a CSR array -> (values, offsets) of the sub-batch (vectorised gather)"""
    offsets = np.asarray(offsets, dtype=np.int64)
    idx = np.asarray(idx, dtype=np.int64)
    lens = offsets[idx + 1] - offsets[idx]
    sub_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(lens, out=sub_off[1:])
    total = int(sub_off[-1])
    src = np.repeat(offsets[idx] - sub_off[:-1], lens) + np.arange(total, dtype=np.int64)
    out = np.asarray(values)[src] if total else np.zeros(0, dtype=np.asarray(values).dtype)
    return np.ascontiguousarray(out), sub_off


def put_csr(dst_values, dst_offsets, idx, sub_values, sub_offsets):
    """inverse of take_csr for per-position outputs (dot-bracket bytes): rows go back to `idx`"""
    dst_offsets = np.asarray(dst_offsets, dtype=np.int64)
    idx = np.asarray(idx, dtype=np.int64)
    lens = np.diff(sub_offsets)
    total = int(sub_offsets[-1])
    if total:
        dst = np.repeat(dst_offsets[idx] - sub_offsets[:-1], lens) + np.arange(total, dtype=np.int64)
        dst_values[dst] = sub_values[:total]


def contiguous_plan(lengths, world, exponent=2.0):
    """Cut the batch into `world` CONTIGUOUS ranges of nearly equal cost (sum of N**exponent): for batches of
    many short sequences, where any range is a fair sample of the lengths -- the ranges are views of the CSR
    arrays, nothing is gathered.  Returns world + 1 boundaries."""
    cost = np.cumsum(np.asarray(lengths, dtype=np.float64) ** exponent)
    n = len(cost)
    if n == 0:
        return np.zeros(world + 1, dtype=np.int64)
    cuts = np.searchsorted(cost, cost[-1] * np.arange(1, world) / world, side="right")
    return np.concatenate(([0], np.minimum(cuts, n), [n])).astype(np.int64)


class MultiGPU:
    """The fast lane (`byseq pl=1` shape) over several GPUs of one box from one process: one host thread per
    GPU (ctypes releases the GIL during the library calls), results in input order."""

    def __init__(self, devices=None, contexts=None):
        from . import _lib
        self.own = contexts is None
        if contexts is not None:
            self.ctx = list(contexts)
            self.devices = [c.device for c in self.ctx]
            return
        n = _lib.load().sqrn_device_count()
        if n < 1:
            raise _lib.SqrnError("no usable CUDA device (squarna_b200 has no CPU fallback)")
        self.devices = list(range(n)) if devices is None else list(devices)
        self.ctx = [_lib.Context(d) for d in self.devices]

    def close(self):
        if self.own:
            for c in self.ctx:
                c.close()

    def fast_predict(self, paramset, symbols, offsets):
        """same contract as Context.fast_predict.  Short sequences (all <= 320 nt: warp teams) are cut into
        contiguous ranges of equal cost; anything longer is dealt by length ** 3 (shard_plan)."""
        offsets = np.asarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        lens = np.diff(offsets)
        world = len(self.ctx)
        dbn = np.empty(max(int(offsets[-1]), 1), dtype=np.uint8)
        scores = np.empty((max(n, 1), 3), dtype=np.float64)
        nst = np.empty(max(n, 1), dtype=np.int32)
        errors = []
        contiguous = n > 0 and int(lens.max()) <= 320
        if contiguous:
            cuts = contiguous_plan(lens, world)
            jobs = [(int(cuts[q]), int(cuts[q + 1])) for q in range(world)]
        else:
            jobs = shard_plan(lens, world, 3.0)

        def work(ctx, job):
            try:
                if contiguous:
                    lo, hi = job
                    if hi <= lo:
                        return
                    t0, t1 = int(offsets[lo]), int(offsets[hi])
                    sub_sym = symbols[t0:t1] if t1 > t0 else np.zeros(1, np.uint8)
                    d, sc, ns = ctx.fast_predict(paramset, np.ascontiguousarray(sub_sym), offsets[lo:hi + 1] - t0)
                    dbn[t0:t1] = d
                    scores[lo:hi] = sc
                    nst[lo:hi] = ns
                    return
                idx = job
                if not len(idx):
                    return
                sub_sym, sub_off = take_csr(symbols, offsets, idx)
                d, sc, ns = ctx.fast_predict(paramset, sub_sym if len(sub_sym) else np.zeros(1, np.uint8), sub_off)
                put_csr(dbn, offsets, idx, d, sub_off)
                scores[idx] = sc
                nst[idx] = ns
            except Exception as e:                     # surfaced in the caller's thread
                errors.append(e)

        threads = [threading.Thread(target=work, args=(c, job)) for c, job in zip(self.ctx, jobs)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        return dbn[:int(offsets[-1])], scores[:n], nst[:n]


# ---------------------------------------------------------------- one process per GPU
def rank_shard(lengths, rank, world, exponent=2.0):
    """the queue of `rank` under torch.distributed (every rank computes the same plan)"""
    return shard_plan(lengths, world, exponent)[rank]


def gather_to_root(idx, per_seq, per_pos, sub_offsets, offsets, rank, world, dist=None):
    """Rank 0 receives every rank's results and puts them back in input order.
    per_seq: dict name -> array with one row per sequence of this rank's shard;
    per_pos: dict name -> flat per-position array (CSR over sub_offsets).
    Returns (per_seq_full, per_pos_full) on rank 0 and (None, None) elsewhere.  The exchange is a
    host-side gather of finished results (gloo or nccl object gather): it is not on the data path."""
    if world == 1:
        parts = [(idx, per_seq, per_pos, sub_offsets)]
    else:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((idx, per_seq, per_pos, sub_offsets), parts, dst=0)
        if rank != 0:
            return None, None
    offsets = np.asarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    seq_full, pos_full = {}, {}
    for pidx, pseq, ppos, poff in parts:
        for name, arr in pseq.items():
            arr = np.asarray(arr)
            if name not in seq_full:
                seq_full[name] = np.zeros((n,) + arr.shape[1:], dtype=arr.dtype)
            seq_full[name][pidx] = arr
        for name, arr in ppos.items():
            arr = np.asarray(arr)
            if name not in pos_full:
                pos_full[name] = np.zeros(int(offsets[-1]), dtype=arr.dtype)
            put_csr(pos_full[name], offsets, pidx, arr, poff)
    return seq_full, pos_full

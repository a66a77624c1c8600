#!/usr/bin/env python
"""bench.py -- throughput of the greedy hot path on BASELINE.json's config 2:
synthetic random RNAs, length U{60..200}, `byseq pl=1 c=fastest.conf`.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--seqs S] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of S sequences per GPU
(default 1 000 000, the configuration the metric is quoted on).  For N > 1 the
driver launches one process per GPU with torchrun; sequences are independent, so
ranks shard them with no data-path collective (weak scaling: S per GPU).

Printed by rank 0 as ONE JSON line:
  value    sequences/s with inputs resident in HBM (device-pointer C-ABI call)
  e2e      same through the host-buffer C-ABI call: pinned host -> device copies
           and device -> host reads inside the timed region
  roofline algorithmic bytes per launch / measured kernel time vs measured HBM peak
  cpu_baseline  the oracle port (oracle/sqrn_oracle.c) on the box's host cores,
           bounded sample of the same workload
`--impl reference` times only the CPU oracle port (rank 0; other ranks exit).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261017
FASTEST = dict(algorithms={"G"}, bpp=0.0, bpweights={"GC": 3.25, "AU": 1.25, "GU": -1.25},
               suboptmax=1.0, suboptmin=1.0, suboptsteps=1.0, minlen=4.0, minbpscore=7.0,
               minfinscorefactor=1.25, distcoef=0.09, bracketweight=-2.0, orderpenalty=1.0,
               loopbonus=0.125, maxstemnum=1e6)          # the reference's fastest.conf
WORKLOAD = "config2: synthetic random RNA, len U{60..200}, byseq pl=1 c=fastest.conf"


def make_batch(n, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(60, 201, size=n, dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    sym = np.frombuffer(b"ACGU", dtype=np.uint8)[rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)]
    return np.ascontiguousarray(sym), off, lens


def algorithmic_bytes(lens):
    """ceil(N/4) + 8 + S*(N + 32) + N per sequence, S = 1 structure"""
    return int(((lens + 3) // 4 + 8 + (lens + 32) + lens).sum())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_oracle_rate(sym, off, lens, sample, threads):
    """oracle port on `threads` host threads over the first `sample` sequences"""
    from oracle import oracle as O
    O.lib()
    s_off = off[:sample + 1]
    s_sym = sym[:int(s_off[-1])]
    t0 = time.perf_counter()
    O.predict_batch_simple(s_sym, s_off, [FASTEST], poollim=1, nthreads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, float((lens[:sample].astype(np.float64) ** 2).sum()) / dt, dt


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = max(1000, min(args.seqs, 3000 * threads))
    sym, off, lens = make_batch(sample, SEED)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(sym, off, lens, min(sample, 200 * threads), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_rate(sym, off, lens, sample, threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    line = {"impl": "reference", "metric": "sequences/sec (SQRNdbnseq greedy, byseq pl=1 fastest.conf)",
            "value": val, "unit": "seq/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "nt2_per_s": float((lens.astype(np.float64) ** 2).sum()) / dt,
            "config": {"workload": WORKLOAD, "seqs_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "seq/s", "cores": threads, "kind": "port",
                             "sample": "first %d sequences of the workload per step, oracle/sqrn_oracle.c "
                                       "(plain-C restatement of the Python reference), %d threads" % (sample, threads)},
            "e2e": {"value": val, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_near_gpu(index):
    """Multi-GPU runs: keep this rank's host threads (and, by first touch, its pinned buffers) on the CPUs NVML
    names as local to its GPU, so that the H2D / D2H traffic of the end-to-end leg does not cross sockets.
    Best effort: any failure leaves the affinity alone.  SQRN_BENCH_NO_BIND=1 switches it off."""
    if os.environ.get("SQRN_BENCH_NO_BIND"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if not all(x.isdigit() for x in ids) or index >= len(ids):
                return None
            index = int(ids[index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in (64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1)
                      if c in allowed)
        if len(cpus) >= 4 and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--seqs", type=int, default=1_000_000, help="sequences per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from squarna_b200 import _lib
    from squarna_b200._abi import ParamSet

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    bound = bind_near_gpu(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    ps = ParamSet.from_dict(FASTEST)

    n = args.seqs
    sym, off, lens = make_batch(n, SEED + rank)
    total, max_len = int(off[-1]), int(lens.max())
    # pinned host buffers (e2e leg) and resident device buffers (value leg)
    h_sym = torch.from_numpy(sym).pin_memory()
    h_off = torch.from_numpy(off).pin_memory()
    h_dbn = torch.empty(total, dtype=torch.uint8).pin_memory()
    h_sc = torch.empty(n * 3, dtype=torch.float64).pin_memory()
    h_ns = torch.empty(n, dtype=torch.int32).pin_memory()
    with torch.cuda.stream(stream):
        d_sym = h_sym.cuda(non_blocking=True)
        d_off = h_off.cuda(non_blocking=True)
        d_dbn = torch.empty(total, dtype=torch.uint8, device="cuda")
        d_sc = torch.empty(n * 3, dtype=torch.float64, device="cuda")
        d_ns = torch.empty(n, dtype=torch.int32, device="cuda")
    stream.synchronize()

    def step_device():
        ctx.fast_predict_device(ps, n, total, max_len, d_off.data_ptr(), d_sym.data_ptr(), d_dbn.data_ptr(),
                                d_sc.data_ptr(), d_ns.data_ptr())

    def step_host():
        rc = ctx.L.sqrn_fast_predict_host(ctx.h, ps, n, h_off.data_ptr(), h_sym.data_ptr(), h_dbn.data_ptr(),
                                          h_sc.data_ptr(), h_ns.data_ptr())
        ctx._check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between barriers; device time by CUDA events on the launch stream, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        stream.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    for _ in range(max(args.warmup, 3)):
        step_device()
    stream.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, _ = timed(step_device, args.steps)
    # kernel time per step: one k_fast launch (the k_fast_rescan launch behind it finds an empty overflow list)
    kern_ms = dev_ms / args.steps
    launches = args.steps * 2          # per step of the timed (device-resident) leg: k_fast<224> + k_fast_rescan<224> behind it
    for _ in range(2):
        step_host()
    _, e2e_wall_ms = timed(step_host, args.steps)
    e2e_stats = ctx.stats()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # sanity: the e2e leg and the device leg produced the same structures
    assert np.array_equal(h_dbn.numpy(), d_dbn.cpu().numpy()), "device and host legs disagree"

    seqs_all = n * world
    value = seqs_all / (dev_ms / args.steps / 1e3)
    nt2 = float((lens.astype(np.float64) ** 2).sum()) * world / (dev_ms / args.steps / 1e3)
    e2e = seqs_all / (e2e_wall_ms / args.steps / 1e3)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        abytes = algorithmic_bytes(lens)
        # DRAM traffic of one launch from the committed ncu capture of this same configuration
        traffic, ncu_note = None, {}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_latest.json")))
            if tr.get("seqs"):
                # per launch of THIS run: the capture's bytes scaled by the sequence count when the sizes differ
                traffic = int((tr["dram_bytes_read"] + tr["dram_bytes_write"]) * (n / float(tr["seqs"])))
                ncu_note = {"issue_slots_busy_pct": tr.get("issue_slots_busy_pct"), "ipc_active": tr.get("ipc_active"),
                            "icc_hit_rate_pct": tr.get("icc_hit_rate_pct"),
                            "warp_instructions_per_sequence": tr.get("warp_instructions_per_sequence"),
                            "traffic_source": tr.get("source"), "traffic_capture_seqs": tr["seqs"]}
        except Exception:
            pass
        achieved = abytes / (kern_ms / 1e3) / 1e9
        threads = os.cpu_count() or 1
        sample = min(n, 50000 * threads)          # ~10 s of host work at ~4 k seq/s per core
        # the host-core baseline is a rank-0, N = 1 leg (the reference arm reports it at every N)
        skip_cpu = args.no_cpu or world > 1
        cpu_rate, cpu_nt2, cpu_dt = (0.0, 0.0, 0.0) if skip_cpu else cpu_oracle_rate(sym, off, lens, sample, threads)
        h2d = int(sym.nbytes + off.nbytes)
        d2h = int(total + n * 3 * 8 + n * 4 + n)
        line = {"metric": "sequences/sec (SQRNdbnseq greedy, byseq pl=1 fastest.conf)", "value": value, "unit": "seq/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "nt2_per_s": nt2,
                "config": {"workload": WORKLOAD, "seqs_per_gpu_per_step": n, "total_nt_per_gpu": total,
                           "l2_policy": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2; no flush" % ((h2d + d2h) / 1e6),
                           "sharding": "independent sequences per rank, no collective",
                           "host_cpus_bound_to_gpu_locality": bound},
                "e2e": {"value": e2e, "unit": "seq/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_wall_ms / args.steps,
                        "kernel_ms_in_step": e2e_stats["kernel_ms"], "launches_per_step": e2e_stats["launches"],
                        "pipeline": "chunked: H2D, kernel and D2H of neighbouring chunks overlap on 4 streams"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                             "kernel": "k_fast<224>", "algorithmic_bytes_per_launch": abytes, "kernel_ms": kern_ms,
                             "note": "issue-bound integer path, not HBM-bound: the kernel keeps ~%s %% of the issue slots busy "
                                     "(profiles/), its DRAM traffic is about the algorithmic bytes"
                                     % ncu_note.get("issue_slots_busy_pct", "80"), **ncu_note},
                "cpu_baseline": {"value": cpu_rate, "unit": "seq/s", "cores": threads, "kind": "port",
                                 "nt2_per_s": cpu_nt2,
                                 "sample": ("not run at N > 1 (see the N = 1 line and the reference arm)" if world > 1 else
                                            "first %d sequences of rank 0's batch, %.1f s, oracle/sqrn_oracle.c on %d threads"
                                            % (sample, cpu_dt, threads))},
                "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

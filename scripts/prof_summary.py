#!/usr/bin/env python
"""Summarise an ncu report (read here, after gpurun brought it back): issue / stall / instruction-cache
metrics, executed code footprint, executed static instructions per function, hot source lines.
usage: prof_summary.py report.ncu-rep [n_seqs] [top]"""
import bisect, csv, io, re, subprocess, sys
rep = sys.argv[1]; nseq = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
def f(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
print("kernel:", d.get("Kernel Name"), " duration ms:", f(d.get("gpu__time_duration.sum", "0")) / 1e6)
for k in ("smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__instruction_throughput.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "sm__icc_request_hit_rate.pct", "sm__icc_requests.sum", "gcc__cache_requests_type_instruction.sum",
          "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "gcc__average_cache_request_hit_rate.pct",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
          "launch__occupancy_limit_registers", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
    print("  %-90s %s" % (k, d.get(k)))
st = [(k, f(v)) for k, v in d.items() if re.search(r"smsp__average_warps_issue_stalled_.*_per_issue_active", k)]
for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
    print("  stall %-40s %.3f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; fpath = ""; per = {}; seen = {}; lines = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if len(r) >= 12 and r[0] not in ("", "Line No"):
        try:
            cur = (fpath, int(r[0])); lines.append((int(r[6]), int(r[7]), int(r[8]), fpath, r[0], r[1].strip()))
        except ValueError: cur = None
        continue
    if len(r) >= 12 and r[0] == "" and r[2].startswith("0x") and cur:
        n = int(r[7]); seen[int(r[2], 16)] = n
        if n > 0: a = per.setdefault(cur, [0, 0]); a[0] += 1; a[1] += n
ex = [a for a, n in seen.items() if n > 0]
print("SASS instructions: %d static, %d executed, executed 128-B lines: %d (%.1f KB)" % (len(seen), len(ex), len(set(a // 128 for a in ex)), len(set(a // 128 for a in ex)) / 8.0))
tot_i = sum(n for n in seen.values())
if nseq: print("warp instructions per sequence: %.0f" % (tot_i / nseq))
dev = open(__file__.rsplit("/", 2)[0] + "/squarna_b200/csrc/sqrn_device.cuh").read().split("\n")
funcs = [(i, re.search(r"(\w+)\s*\(", l).group(1)) for i, l in enumerate(dev, 1) if l.startswith("__device__") and re.search(r"(\w+)\s*\(", l)]
starts = [x[0] for x in funcs]; agg = {}
for (fn, ln), (s_, dy) in per.items():
    name = fn
    if fn == "sqrn_device.cuh":
        k = bisect.bisect_right(starts, ln) - 1; name = funcs[k][1] if k >= 0 else "?"
    a = agg.setdefault(name, [0, 0]); a[0] += s_; a[1] += dy
tot = sum(v[1] for v in agg.values()) or 1
print("executed static instructions / share of dynamic instructions, per function:")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]: print("  %-28s %5d  %5.1f %%" % (k, v[0], 100.0 * v[1] / tot))
lines.sort(reverse=True); ts = sum(x[0] for x in lines) or 1; ti = sum(x[1] for x in lines) or 1
print("%6s %6s %5s  %s" % ("smp%", "ins%", "thr", "line"))
for s_, i, t, fn, ln, text in lines[:top]:
    print("%6.2f %6.2f %5.1f  %s:%s  %s" % (100.0 * s_ / ts, 100.0 * i / ti, t / max(i, 1), fn, ln, text[:100]))

#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: ncu_lines.py src_cs.csv [top]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []; hdr = None; fpath = ""
tot_s = tot_i = 0
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10: continue
    if r[0] != "":            # a source line row carries the aggregate of its SASS
        try:
            s = int(r[6]); i = int(r[7]); t = int(r[8])
        except ValueError:
            continue
        rows.append((s, i, t, fpath, r[0], r[1].strip()))
        tot_s += s; tot_i += i
rows.sort(reverse=True)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
print("%6s %6s %12s %5s  %s" % ("smp%", "ins%", "instr", "thr", "line"))
for s, i, t, f, ln, src in rows[:top]:
    print("%6.2f %6.2f %12d %5.1f  %s:%s  %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), i, t / max(i, 1), f, ln, src[:110]))

#!/usr/bin/env python
"""Static SASS instruction count per CUDA source line for one kernel.
usage: sass_lines.py <lib.so> <kernel-substring> [top]"""
import re, subprocess, sys, tempfile, os, glob
lib, pat = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
txt = subprocess.run(["nvdisasm", "--print-line-info"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
cnt = {}; cur = None; infn = False; total = 0
for line in txt.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m: infn = pat in m.group(1); cur = None; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
        cnt[cur] = cnt.get(cur, 0) + 1; total += 1
print("total", total)
src = {}
for (f, l), c in sorted(cnt.items(), key=lambda x: -x[1])[:top] if cnt else []:
    if f not in src:
        for base in ("squarna_b200/csrc/", "include/", ""):
            try: src[f] = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)).read().split("\n"); break
            except Exception: src[f] = None
    text = src[f][l - 1].strip()[:100] if src.get(f) and l <= len(src[f]) else ""
    print("%5d  %s:%d  %s" % (c, f, l, text))

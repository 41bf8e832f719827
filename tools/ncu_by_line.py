#!/usr/bin/env python
"""Attribute an ncu SASS source page to CUDA source lines using nvdisasm line info.
usage: ncu_by_line.py <sass.csv from `ncu --page source --csv --print-source sass`> <nvdisasm -g -c output> <mangled kernel substring> [source dir]"""
import csv, re, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; rows = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
lines = []   # source line of each SASS instruction of the kernel, in order
cur = None; inside = False
for l in open(sys.argv[2]):
    if l.startswith('.text.'):
        inside = sys.argv[3] in l
        continue
    if not inside: continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(cur)
print("sass rows", len(rows), "disasm instructions", len(lines), file=sys.stderr)
n = min(len(rows), len(lines))
inst = defaultdict(int); samp = defaultdict(int); thr = defaultdict(int)
for r, ln in zip(rows[:n], lines[:n]):
    inst[ln] += int(r[ix['Instructions Executed']]); samp[ln] += int(r[ix['# Samples']]); thr[ln] += int(r[ix['Thread Instructions Executed']])
ti = sum(inst.values()); ts = sum(samp.values())
import os
srcdir = sys.argv[4] if len(sys.argv) > 4 else None       # directory holding the .cu/.cuh sources
cache = {}
def text(key):
    if not srcdir or key is None: return ''
    f, ln = key
    if f not in cache:
        try: cache[f] = open(os.path.join(srcdir, f)).read().splitlines()
        except OSError: cache[f] = []
    return cache[f][ln - 1].strip()[:100] if 0 < ln <= len(cache[f]) else ''
print("warp instructions %d, samples %d, average active lanes %.1f" % (ti, ts, sum(thr.values()) / max(ti, 1)))
for key in sorted(inst, key=lambda k: (k is None, k)):
    if inst[key] / ti > 0.003 or samp[key] / ts > 0.003:
        f, ln = key if key else ('?', 0)
        print(f"{f:14s}{ln:5d} inst {inst[key]/ti*100:5.1f}%  samples {samp[key]/ts*100:5.1f}%  lanes {thr[key]/max(inst[key],1):5.1f} | {text(key)}")

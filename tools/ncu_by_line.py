#!/usr/bin/env python
"""Attribute an ncu SASS source page to CUDA source lines using nvdisasm line info.
usage: ncu_by_line.py <sass.csv from `ncu --page source --csv --print-source sass`> <nvdisasm -g -c output> <mangled kernel substring>"""
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
    m = re.search(r'//## File ".*", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l): lines.append(cur)
print("sass rows", len(rows), "disasm instructions", len(lines), file=sys.stderr)
n = min(len(rows), len(lines))
inst = defaultdict(int); samp = defaultdict(int); thr = defaultdict(int)
for r, ln in zip(rows[:n], lines[:n]):
    inst[ln] += int(r[ix['Instructions Executed']]); samp[ln] += int(r[ix['# Samples']]); thr[ln] += int(r[ix['Thread Instructions Executed']])
ti = sum(inst.values()); ts = sum(samp.values())
src = open(sys.argv[4]).read().splitlines() if len(sys.argv) > 4 else None
for ln in sorted(inst):
    if inst[ln] / ti > 0.003 or samp[ln] / ts > 0.003:
        print(f"{ln:5d} inst {inst[ln]/ti*100:5.1f}%  samples {samp[ln]/ts*100:5.1f}%  lanes {thr[ln]/max(inst[ln],1):5.1f} | {(src[ln-1].strip()[:100] if src else '')}")

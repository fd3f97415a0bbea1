#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line / function.

usage: ncu_by_line.py <sass_page.csv> <nvdisasm -g listing> <kernel mangled name> <header path>
Joins per-instruction 'Instructions Executed' / stall samples (ncu --page source
--csv) with nvdisasm's '//## File "...", line N' annotations.
"""
import csv
import re
import sys
from collections import defaultdict

sass_csv, listing, kernel, header = sys.argv[1:5]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ii, isamp, ithr = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Thread Instructions Executed')
insts = [(int(r[ia], 16), int(r[ii]), int(r[isamp]), int(r[ithr])) for r in rows[2:] if len(r) > ii and r[ia].startswith('0x')]
base = insts[0][0]
by_off = {a - base: (n, s, t) for a, n, s, t in insts}

# function line ranges in the header
src = open(header).read().split('\n')
funcs = []
for ln, text in enumerate(src, 1):
    mm = re.match(r'\s*FB_(?:MEM|DEV)\s+[\w\s\*]+?\b(\w+)\(', text)
    if mm:
        funcs.append((ln, mm.group(1)))
def func_of(line):
    name = '?'
    for ln, fn in funcs:
        if ln <= line:
            name = fn
    return name

cur = None
infn = False
line_tot = defaultdict(lambda: [0, 0, 0])
fn_tot = defaultdict(lambda: [0, 0, 0])
for text in open(listing):
    if text.startswith('.text.'):
        infn = text.strip().rstrip(':') == '.text.' + kernel
        continue
    if not infn:
        continue
    mm = re.search(r'//## File "([^"]+)", line (\d+)', text)
    if mm:
        cur = (mm.group(1).split('/')[-1], int(mm.group(2)))
        continue
    mm = re.match(r'\s*/\*([0-9a-f]{4,})\*/', text)
    if mm and cur:
        off = int(mm.group(1), 16)
        if off in by_off:
            n, s, t = by_off[off]
            key = cur
            line_tot[key][0] += n; line_tot[key][1] += s; line_tot[key][2] += t
            fn = func_of(cur[1]) if cur[0].endswith('fb_device.h') else cur[0]
            fn_tot[fn][0] += n; fn_tot[fn][1] += s; fn_tot[fn][2] += t
tot = sum(v[0] for v in fn_tot.values()); tots = sum(v[1] for v in fn_tot.values())
print(f'total warp-insts {tot}  samples {tots}')
print('--- by function: inst%  sample%  avg-threads')
for fn, (n, s, t) in sorted(fn_tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{fn:28s} {100*n/tot:6.2f}% {100*s/max(1,tots):6.2f}%  {t/max(1,n):5.1f}')
print('--- top lines by samples')
for (f, ln), (n, s, t) in sorted(line_tot.items(), key=lambda kv: -kv[1][1])[:45]:
    text = src[ln-1].strip()[:90] if f.endswith('fb_device.h') else ''
    print(f'{f}:{ln:5d} inst {100*n/tot:5.2f}% samp {100*s/max(1,tots):5.2f}% thr {t/max(1,n):4.1f} | {text}')

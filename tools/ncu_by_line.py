#!/usr/bin/env python
"""Aggregate an ncu SASS source page by CUDA source line / function.

usage: ncu_by_line.py <sass_page.csv> <nvdisasm -g listing> <kernel mangled name> <csrc dir>
Joins per-instruction 'Instructions Executed' / stall samples (ncu --page source
--csv) with nvdisasm's '//## File "...", line N' annotations, attributes every
line to the enclosing FB_MEM/FB_DEV function of the header it lives in, and
prints the stall-reason mix of the kernel and of the heaviest lines.
"""
import csv
import os
import re
import sys
from collections import defaultdict

sass_csv, listing, kernel, csrc = sys.argv[1:5]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, ii, isamp, ithr = (hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples'),
                       hdr.index('Thread Instructions Executed'))
isrc = hdr.index('Source')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
insts = []
for r in rows[2:]:
    if len(r) > ii and r[ia].startswith('0x'):
        insts.append((int(r[ia], 16), int(r[ii]), int(r[isamp]), int(r[ithr]),
                      [int(r[i] or 0) for i, _ in stall_cols], r[isrc].split()[0] if r[isrc].split() else '?'))
base = insts[0][0]
by_off = {a - base: (n, s, t, st, op) for a, n, s, t, st, op in insts}

sources, funcs = {}, {}
for name in os.listdir(csrc):
    if not name.endswith(('.h', '.cu')):
        continue
    text = open(os.path.join(csrc, name)).read().split('\n')
    sources[name] = text
    fl = []
    for ln, line in enumerate(text, 1):
        mm = re.match(r'\s*FB_(?:MEM|DEV)\s+[\w\s\*&]+?\b(\w+)\(', line)
        if mm:
            fl.append((ln, mm.group(1)))
    funcs[name] = fl


def func_of(fname, line):
    name = fname
    for ln, fn in funcs.get(fname, []):
        if ln <= line:
            name = fn
    return name


cur, infn = None, False
line_tot = defaultdict(lambda: [0, 0, 0])
fn_tot = defaultdict(lambda: [0, 0, 0])
line_stall = defaultdict(lambda: [0]*len(stall_cols))
op_tot = defaultdict(lambda: [0, 0])
tot_stall = [0]*len(stall_cols)
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
            n, s, t, st, op = by_off[off]
            line_tot[cur][0] += n; line_tot[cur][1] += s; line_tot[cur][2] += t
            fn = func_of(*cur)
            fn_tot[fn][0] += n; fn_tot[fn][1] += s; fn_tot[fn][2] += t
            op_tot[op.split('.')[0]][0] += n; op_tot[op.split('.')[0]][1] += s
            for k, v in enumerate(st):
                line_stall[cur][k] += v
                tot_stall[k] += v
tot = sum(v[0] for v in fn_tot.values())
tots = sum(v[1] for v in fn_tot.values())
print(f'total warp-insts {tot}  samples {tots}')
print('--- stall reasons (share of samples)')
ssum = max(1, sum(tot_stall))
print('  '.join(f'{h[6:]} {100*v/ssum:.1f}%' for (_, h), v in sorted(zip(stall_cols, tot_stall), key=lambda kv: -kv[1]) if v*100 > ssum))
print('--- by opcode: inst%  sample%')
for op, (n, s) in sorted(op_tot.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f'{op:10s} {100*n/tot:6.2f}% {100*s/max(1,tots):6.2f}%')
print('--- by function: inst%  sample%  avg-threads')
for fn, (n, s, t) in sorted(fn_tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{fn:28s} {100*n/tot:6.2f}% {100*s/max(1,tots):6.2f}%  {t/max(1,n):5.1f}')
print('--- top lines by samples')
for (f, ln), (n, s, t) in sorted(line_tot.items(), key=lambda kv: -kv[1][int(os.environ.get("BY_INST","0")) ^ 1])[:int(os.environ.get("TOPN","45"))]:
    text = sources[f][ln-1].strip()[:80] if f in sources and ln <= len(sources[f]) else ''
    st = line_stall[(f, ln)]
    top = sorted(zip(st, [h[6:] for _, h in stall_cols]), reverse=True)[:2]
    tops = ','.join(f'{nm}:{100*v/max(1,sum(st)):.0f}' for v, nm in top)
    print(f'{f}:{ln:5d} inst {100*n/tot:5.2f}% samp {100*s/max(1,tots):5.2f}% [{tops}] | {text}')

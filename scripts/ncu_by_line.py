#!/usr/bin/env python
"""Aggregate an ncu source-page CSV (SASS rows) by CUDA source line using nvdisasm -g line info of the same cubin.
usage: ncu_by_line.py <sass_with_lineinfo.txt> <mangled kernel name> <ncu_source.csv> [top]"""
import csv, re, sys, collections
sass, kern, ncsv = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.' + kern + ':')][0]
cur = None; per_inst = []
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .text.') or l.startswith('\t.section'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        per_inst.append(cur)
rows = list(csv.reader(open(ncsv)))
hdr = rows[1]; iN = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
data = []
for r in rows[2:]:
    if len(r) <= iN or r[iN] == 'Instructions Executed' or r[0] == 'Kernel Name':
        break
    data.append(r)
print('sass insts', len(per_inst), 'csv rows', len(data))
agg = collections.defaultdict(lambda: [0, 0])
for li, r in zip(per_inst, data):
    agg[li][0] += int(r[iN]); agg[li][1] += int(r[iS])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = k if k else ('?', 0)
    text = ''
    if f.endswith('.cu') or f.endswith('.cuh'):
        import os
        pth = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'ecwam_b200', 'csrc', f)
        if pth not in src and os.path.exists(pth): src[pth] = open(pth).read().split('\n')
        if pth in src and 0 < ln <= len(src[pth]): text = src[pth][ln - 1].strip()[:90]
    print(f'{f}:{ln:5d} inst {v[0]/ti*100:5.2f}% samp {v[1]/ts*100:5.2f}%  {text}')

#!/usr/bin/env python
"""Per-region instruction / stall-sample / shared-wavefront totals from `ncu --page source --csv` output.
usage: src_hot.py file.csv [bucket_size_in_instructions]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iA, iS, iN, iI = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
iW = hdr.index('L1 Wavefronts Shared'); iWi = hdr.index('L1 Wavefronts Shared Ideal')
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
body = [r for r in rows[2:] if len(r) > max(iI, iW, iWi) and r[iI].isdigit()]
tot_i = sum(int(r[iI]) for r in body); tot_s = sum(int(r[iN]) for r in body)
print('total inst', tot_i, 'samples', tot_s, 'n sass', len(body))
for b in range(0, len(body), B):
    ch = body[b:b + B]
    ins = sum(int(r[iI]) for r in ch); sm = sum(int(r[iN]) for r in ch)
    wf = sum(int(r[iW]) for r in ch); wfi = sum(int(r[iWi]) for r in ch)
    tags = [r[iS].split()[0] if not r[iS].strip().startswith('@') else r[iS].split()[1] for r in ch]
    marks = [t for t in tags if t.startswith(('BAR', 'LDGSTS', 'ATOMS', 'REDG', 'STG', 'LDG', 'REDUX', 'EXIT', 'ATOMG'))]
    from collections import Counter
    c = Counter(m.split('.')[0] for m in marks)
    print(f'{b:5d} inst {ins/tot_i*100:5.1f}%  samples {sm/max(tot_s,1)*100:5.1f}%  wf {wf:9d} ideal {wfi:9d}  {dict(c)}')

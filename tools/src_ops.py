#!/usr/bin/env python
"""Opcode-level totals (instructions, shared wavefronts, stall samples) from `ncu --page source --csv` of ONE kernel.
usage: ncu -i rep --page source --csv --kernel-name regex:NAME > f.csv ; src_ops.py f.csv [launches_in_file]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iW, iWi, iN = (hdr.index(k) for k in ('Source', 'Instructions Executed', 'L1 Wavefronts Shared', 'L1 Wavefronts Shared Ideal', '# Samples'))
body = [r for r in rows[2:] if len(r) > iWi and r[iI].isdigit()]
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1
body = body[:len(body) // L]
tot = sum(int(r[iI]) for r in body)
print('n sass', len(body), 'total inst %.1fM' % (tot / 1e6), 'samples', sum(int(r[iN]) for r in body))
agg = defaultdict(lambda: [0, 0, 0, 0])
for r in body:
    t = r[iS].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')
    op = op[0] + ('.' + op[-1] if op[0] in ('LDS', 'STS', 'LDG', 'STG') and op[-1] in ('64', '128') else '')
    a = agg[op]
    a[0] += int(r[iI]); a[1] += int(r[iW]); a[2] += int(r[iWi]); a[3] += int(r[iN])
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
    print(f'{op:10s} inst {a[0]/1e6:8.2f}M ({a[0]/tot*100:4.1f}%)  wf {a[1]/1e6:7.2f}M ideal {a[2]/1e6:7.2f}M  samples {a[3]}')

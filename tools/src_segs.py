#!/usr/bin/env python
"""Per barrier-delimited segment: instructions per warp, stall samples, shared wavefronts (ncu --page source --csv)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iN, iW = (hdr.index(k) for k in ('Source', 'Instructions Executed', '# Samples', 'L1 Wavefronts Shared'))
body = [r for r in rows[2:] if len(r) > iW and r[iI].isdigit()]
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1
warps = float(sys.argv[3]) if len(sys.argv) > 3 else 32768.0
body = body[:len(body) // L]
tot = sum(int(r[iI]) for r in body)
print(len(body), 'sass;', tot / 1e6, 'M inst')
seg = acc = accs = accw = start = 0
for k, r in enumerate(body):
    acc += int(r[iI]); accs += int(r[iN]); accw += int(r[iW])
    if 'BAR.SYNC' in r[iS] or k == len(body) - 1:
        print(f'seg {seg} sass {start}-{k}: inst/warp {acc/warps:8.1f}  ({acc/tot*100:4.1f}%) samples {accs} wf {accw/1e6:.2f}M')
        seg += 1; acc = accs = accw = 0; start = k + 1

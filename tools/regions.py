#!/usr/bin/env python
"""usage: regions.py src.csv a:b:name ...   — instruction / sample / smem-wavefront totals of SASS index ranges"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1], errors='replace')))
hdr=rows[1]; iS=hdr.index('Source'); iN=hdr.index('# Samples'); iI=hdr.index('Instructions Executed'); iW=hdr.index('L1 Wavefronts Shared')
body=[r for r in rows[2:] if len(r)>iW and r[iI].isdigit()]
ti=sum(int(r[iI]) for r in body); ts=sum(int(r[iN]) for r in body)
print('total %.1fM inst, %d samples, %d sass'%(ti/1e6, ts, len(body)))
for spec in sys.argv[2:]:
    a,b,name=spec.split(':'); ch=body[int(a):int(b)]
    print(f"{name:18s} inst {sum(int(r[iI]) for r in ch)/1e6:7.1f}M {sum(int(r[iI]) for r in ch)/ti*100:5.1f}%  samples {sum(int(r[iN]) for r in ch)/ts*100:5.1f}%  wf {sum(int(r[iW]) for r in ch)/1e6:6.1f}M")

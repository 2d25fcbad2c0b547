"""hot SASS regions of one kernel from an ncu source-page csv: usage sass_hot.py file.csv"""
import csv, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>=10 and r[0].startswith('0x')]
tot=sum(int(r[5]) for r in rows); ts=sum(int(r[4]) for r in rows)
print('total warp instr',tot,'n sass',len(rows),'samples',ts)
prev=None; start=0; acc=0; samples=0; out=[]
for i,r in enumerate(rows):
    c=int(r[5])
    if prev is not None and abs(c-prev)>0.02*max(c,prev,1):
        out.append((start,i-1,prev,acc,samples)); start=i; acc=0; samples=0
    acc+=c; samples+=int(r[4]); prev=c
out.append((start,len(rows)-1,prev,acc,samples))
for s,e,c,a,sm in out:
    if a>tot*0.01 or sm>ts*0.01:
        ops={}
        for r in rows[s:e+1]:
            op=r[1].split()[0] if not r[1].strip().startswith('@') else r[1].split()[1]
            op=op.split('.')[0]; ops[op]=ops.get(op,0)+1
        top=sorted(ops.items(),key=lambda x:-x[1])[:8]
        print('sass[%d..%d] n=%d exec=%d instr=%.1f%% samples=%.1f%% %s'%(s,e,e-s+1,c,100*a/tot,100*sm/ts,top))

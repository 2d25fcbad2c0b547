"""per-source-line instruction counts of one kernel: usage src_hot.py rep.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+rx],capture_output=True,text=True).stdout
cur=None; rows=[]; seen_fn=set(); first_fn=None
fn=None
for r in csv.reader(out.splitlines()):
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=='Function Name':
        fn=r[1]
        if first_fn is None: first_fn=fn
        continue
    if len(r)>8 and r[0].isdigit() and fn==first_fn:
        try: rows.append((cur,int(r[0]),r[1].strip(),int(r[7]),int(r[6])))
        except ValueError: pass
tot=sum(x[3] for x in rows); ts=sum(x[4] for x in rows)
print('kernel',first_fn,'total warp instr',tot,'samples',ts)
byfile={}
for f,l,s,c,sm in rows: byfile[f]=byfile.get(f,0)+c
print({k:'%.1f%%'%(100*v/tot) for k,v in byfile.items()})
for f,l,s,c,sm in sorted(rows,key=lambda x:-x[3])[:top]:
    print('%5.1f%% instr %5.1f%% samp  %s:%d  %s'%(100*c/tot,100*sm/max(ts,1),f,l,s[:100]))

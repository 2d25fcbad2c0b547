import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# split kernels
kern = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; kern.append(cur); continue
    if cur is not None: cur["rows"].append(r)
for k in kern:
    hdr = k["rows"][0]; body = [r for r in k["rows"][1:] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    tot_s = sum(int(r[ix['# Samples']]) for r in body); tot_i = sum(int(r[ix['Instructions Executed']]) for r in body)
    print("==", k["name"][:80], "sass", len(body), "samples", tot_s, "inst", tot_i)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in body:
        for s in stalls: agg[s] += int(r[ix[s]])
    print("  stall totals:", ", ".join(f"{s[6:]} {v/tot_s*100:.1f}%" for s, v in agg.most_common(10)))
    # region buckets of 64 instrs
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    for b in range(0, len(body), B):
        ch = body[b:b+B]
        sm = sum(int(r[ix['# Samples']]) for r in ch); ins = sum(int(r[ix['Instructions Executed']]) for r in ch)
        wf = sum(int(r[ix['L1 Wavefronts Shared']]) for r in ch) if 'L1 Wavefronts Shared' in ix else 0
        ops = collections.Counter(r[ix['Source']].split()[1 if r[ix['Source']].strip().startswith('@') else 0].split('.')[0] for r in ch)
        mark = {o: c for o, c in ops.items() if o in ('BAR','LDGSTS','ATOMS','RED','STG','LDG','LDS','STS','BRA','EXIT','SHFL','LDSM','ATOM','REDG','REDUX')}
        st = collections.Counter()
        for r in ch:
            for s in stalls: st[s] += int(r[ix[s]])
        top = ", ".join(f"{s[6:]} {v/max(sm,1)*100:.0f}%" for s, v in st.most_common(3))
        print(f"  {b:5d} inst {ins/tot_i*100:5.1f}% samp {sm/tot_s*100:5.1f}% wf {wf/1e6:6.2f}M  {mark}  [{top}]")

"""Probe which association/contraction ATen uses for grid_sample unnormalisation on a device (cpu|cuda).
A parity image with H=1 makes the bilinear output equal frac(ix) exactly; each candidate formula is compared
bit for bit.  Result recorded in DESIGN.md (Coordinate arithmetic).  usage: python tools/probe_coords.py cuda"""
import torch, numpy as np, torch.nn.functional as F, sys
dev=sys.argv[1] if len(sys.argv)>1 else 'cpu'
def variants(g, size, align):
    g=g.astype(np.float32); one=np.float32(1); s=np.float32(size)
    out={}
    if align:
        out['cuda']= ((g+one)/np.float32(2))*np.float32(size-1)
        out['cpu']= (g+one)*(np.float32(size-1)/np.float32(2))
    else:
        a=(g+one)
        out['cuda_nofma']=((a*s)-one)/np.float32(2)
        out['cuda_fma']=((a.astype(np.float64)*np.float64(s)-1.0).astype(np.float32))/np.float32(2)
        sf=np.float32(size)/np.float32(2)
        out['cpu_nofma']=(a*sf)-np.float32(0.5)
        out['cpu_fma']=(a.astype(np.float64)*np.float64(sf)-0.5).astype(np.float32)
        out['decomp']=g*sf+np.float32((size-1)/2)
    return out
rng=np.random.default_rng(0)
for size in [7,23,128,150,257,1000,2047]:
  for align in [False,True]:
    M=1<<20
    g=rng.uniform(-1,1,M).astype(np.float32)
    img=(torch.arange(size)%2).float().view(1,1,1,size).to(dev)
    grid=torch.zeros(1,1,M,2); grid[0,0,:,0]=torch.from_numpy(g)
    if align: pass  # H=1, align True: iy=((gy+1)/2)*0=0
    o=F.grid_sample(img,grid.to(dev),mode='bilinear',padding_mode='zeros',align_corners=align)[0,0,0].cpu().numpy()
    res={}
    for name,ix in variants(g,size,align).items():
        x0=np.floor(ix); fr=ix-x0
        inb0=(x0>=0)&(x0<size); inb1=(x0+1>=0)&(x0+1<size)
        # value: img[x0]*(x1-ix) + img[x0+1]*(ix-x0)
        v0=np.where(inb0,(x0%2),0).astype(np.float32); v1=np.where(inb1,((x0+1)%2),0).astype(np.float32)
        pred=v0*((x0+1)-ix).astype(np.float32)+v1*fr.astype(np.float32)
        res[name]=int((pred!=o).sum())
    print(size,align,res)

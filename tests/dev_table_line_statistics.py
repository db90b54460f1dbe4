"""Development analysis (not a test, not product code; the oracle is only the source of the graphs): how many distinct 128-byte
lines of an 8-byte-per-key table the smoothed cells of a warp (32 consecutive cells of a 34 x 34 cell tile in tile order) touch
on the bench frames, for the key layout as it is, with the neighbour bits as the LOW key bits, and for permutations of the keys
by frequency / by popcount — DESIGN.md, "Looked at and not built in round 4".  ncu's L1 tag requests per gather instruction of
raster_kernel<4> (14.6) match the first number (15.0).  python tests/dev_table_line_statistics.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle, build
from pixel_art_remaster_gpu_b200 import synth
build(ref=False); o=Oracle()
W,H=256,224
tot={}
nw=0
for seed in range(4):
    img=synth.snes_stream(1,W,H,first_seed=synth.BASE_SEED+seed)[0]
    out=o.pipeline(img, subdivide=True, want=("graph",))
    g=out["graph"].reshape(H,W).astype(np.uint32)
    # key = node | left.bit2<<8 | left.bit7<<9 | right.bit0<<10 | right.bit5<<11  (linear index neighbours)
    flat=g.reshape(-1)
    left=np.roll(flat,1); right=np.roll(flat,-1)
    key=(flat | ((left>>2)&1)<<8 | ((left>>7)&1)<<9 | (right&1)<<10 | ((right>>5)&1)<<11).reshape(H,W)
    smoothed = (g!=90)
    # tiles of 32x32 with 1 halo: cells 34x34 in tile order idx -> warps of 32 consecutive idx
    for ty in range(0,H,32):
        for tx in range(0,W,32):
            ys=np.clip(np.arange(ty-1,ty+33),0,H-1); xs=np.clip(np.arange(tx-1,tx+33),0,W-1)
            k=key[np.ix_(ys,xs)].reshape(-1); sm=smoothed[np.ix_(ys,xs)].reshape(-1)
            for w0 in range(0,len(k),32):
                kk=k[w0:w0+32]; ss=sm[w0:w0+32]
                if ss.sum()==0: continue
                ks=kk[ss]
                rot=((ks&0xFF)<<4)|(ks>>8)
                for name,val in (("head_now",len(np.unique(ks>>4))),("head_rot",len(np.unique(rot>>4))),("keys",len(np.unique(ks))),("nodes",len(np.unique(ks&0xFF))),("n",len(ks))):
                    tot[name]=tot.get(name,0)+val
                nw+=1
print({k:round(v/nw,2) for k,v in tot.items()}, nw)

# --- orderings
def collect(seeds):
    allk=[]; warps=[]
    for seed in seeds:
        img=synth.snes_stream(1,W,H,first_seed=synth.BASE_SEED+seed)[0]
        g=o.pipeline(img, subdivide=True, want=("graph",))["graph"].reshape(H,W).astype(np.uint32)
        flat=g.reshape(-1); left=np.roll(flat,1); right=np.roll(flat,-1)
        key=(flat | ((left>>2)&1)<<8 | ((left>>7)&1)<<9 | (right&1)<<10 | ((right>>5)&1)<<11).reshape(H,W)
        sm=(g!=90)
        for ty in range(0,H,32):
            for tx in range(0,W,32):
                ys=np.clip(np.arange(ty-1,ty+33),0,H-1); xs=np.clip(np.arange(tx-1,tx+33),0,W-1)
                k=key[np.ix_(ys,xs)].reshape(-1); s=sm[np.ix_(ys,xs)].reshape(-1)
                for w0 in range(0,len(k),32):
                    ks=k[w0:w0+32][s[w0:w0+32]]
                    if len(ks): warps.append(ks); allk.append(ks)
    return np.concatenate(allk), warps
trainK,_=collect(range(100,104))
_,testW=collect(range(0,3))
def evalperm(rank,name):
    print(name, round(np.mean([len(np.unique(rank[w]>>4)) for w in testW]),2))
ident=np.arange(4096); evalperm(ident,"identity")
freq=np.bincount(trainK,minlength=4096); order=np.argsort(-freq,kind='stable'); rank=np.empty(4096,int); rank[order]=np.arange(4096); evalperm(rank,"by frequency (train seeds)")
pc=np.array([bin(k&0xFF).count('1') for k in range(4096)]); order=np.lexsort((ident,-pc)); rank=np.empty(4096,int); rank[order]=np.arange(4096); evalperm(rank,"by popcount desc")
print("distinct keys seen in train:", (freq>0).sum(), "top-64 coverage", freq[np.argsort(-freq)[:64]].sum()/freq.sum(), "top-256", freq[np.argsort(-freq)[:256]].sum()/freq.sum())

"""Development analysis (not a test, not product code; the oracle is only the source of the graphs): ambiguous blocks (both
diagonals survive stage B) and blocks that need the curve-length walks, per pixel of the bench frames — DESIGN.md, round 4, the
sparse form of stage C.  python tests/dev_ambiguous_block_statistics.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle, build
from pixel_art_remaster_gpu_b200 import synth
build(ref=False); o=Oracle()
W,H=256,224
for seed in range(3):
    img=synth.snes_stream(1,W,H,first_seed=synth.BASE_SEED+seed)[0]
    out=o.pipeline(img, subdivide=True, want=("graph_aux","graph"))
    a=out["graph_aux"].reshape(H,W).astype(np.uint32); g=out["graph"].reshape(H,W)
    i1=a[:-1,:-1]; i2=a[1:,:-1]; i3=a[:-1,1:]; i4=a[1:,1:]
    amb=((i1>>2)&1)&((i4>>5)&1)&((i2>>7)&1)&(i3&1)
    pc=lambda x: np.array([bin(v).count('1') for v in range(256)])[x]
    o1=pc(i1&251); o4=pc(i4&223); o3=pc(i3&254); o2=pc(i2&127)
    r1=(o1==1)&(o4==1); r2=~r1&(o3==1)&(o2==1); r3=~r1&~r2&((o1==0)|(o4==0))&(o3!=0)&(o2!=0); r4=~r1&~r2&~r3&((o3==0)|((o2==0)&(o1!=0)&(o4!=0)))
    pend=amb.astype(bool)&~(r1|r2|r3|r4)
    # tiles with any ambiguous
    t=0;n=0
    for ty in range(0,H,32):
        for tx in range(0,W,64):
            n+=1; t+= amb[ty:ty+32,tx:tx+64].any()
    print(seed, 'amb/px', amb.sum()/(W*H), 'pending/px', pend.sum()/(W*H), 'tiles with amb', t, n, 'changed px', (g!=out["graph_aux"].reshape(H,W)).mean())

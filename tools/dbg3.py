import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
from oracle.oracle import Oracle
o = Oracle()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
img = synth.snes_frame(96, 80, synth.BASE_SEED + 1)
want = o.pipeline(img, scale=S, want=("graph", "raster", "poly", "poly_count"))
ctx = par.Remaster(0, 96, 80, 1)
g = torch.from_numpy(want["graph"][None]).cuda()
fr = torch.from_numpy(img[None]).cuda()
for no_tma in (False, True):
    got = ctx.raster(fr, g, S, True, no_tma=no_tma)[0].cpu().numpy()
    bad = np.argwhere((got != want["raster"]).any(-1))
    print("no_tma", no_tma, "mismatches", len(bad))
    from collections import Counter
    print(" sub-pixel (x%S,y%S) histogram:", Counter((int(x) % S, int(y) % S) for y, x in bad).most_common(12))
    for y, x in bad[:6]:
        i, j = x // S, y // S
        n = j * 96 + i
        print((x, y), "cell", (i, j), "sub", (x % S, y % S), "got", got[y, x], "want", want["raster"][y, x], "own", img[j, i][::-1],
              "cnt", want["poly_count"][n], "poly", want["poly"][n][:want["poly_count"][n]].tolist())

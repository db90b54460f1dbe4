import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
from oracle.oracle import Oracle
o = Oracle()
img = synth.snes_frame(96, 80, synth.BASE_SEED + 1)
want = o.pipeline(img, want=("graph", "labels"))
ctx = par.Remaster(0, 96, 80, 1)
g = torch.from_numpy(want["graph"][None]).cuda()
lab = ctx.cc_labels(g)[0].cpu().numpy()
bad = np.argwhere(lab != want["labels"])
print("mismatches", len(bad))
for y, x in bad[:12]:
    print((x, y), "got", lab[y, x], "want", want["labels"][y, x], "node", bin(want["graph"][y, x]))

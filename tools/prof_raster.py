import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
frames = torch.from_numpy(synth.snes_stream(F, 256, 224)).cuda()
ctx = par.Remaster(0, 256, 224, F)
for it in range(3):
    out = ctx.remaster(frames, S, True, want=("rgba", "graph", "graph_aux", "labels"))
    del out
idx = ctx.remaster(frames, S, True, out_format=par.OUT_INDEX8)  # + the palette kernels, the raster kernel writing indices
torch.cuda.synchronize()

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
img = synth.snes_frame(96, 80, 1)
frames = torch.from_numpy(img[None]).cuda()
ctx = par.Remaster(0, 96, 80, 1)
for no_tma in (True, False):
    try:
        aux = ctx.similarity_graph(frames, no_tma=no_tma); torch.cuda.synchronize(); print("K1 ok no_tma", no_tma)
        g = ctx.resolve_crossings(aux, no_tma=no_tma); torch.cuda.synchronize(); print("K2 ok no_tma", no_tma)
        r = ctx.raster(frames, g, 4, True, no_tma=no_tma); torch.cuda.synchronize(); print("K4 ok no_tma", no_tma)
    except Exception as e:
        print("FAIL no_tma", no_tma, e); break

"""One raster launch for ncu (development aid): python tools/prof_k4.py frames scale subdivide(0|1); PAR_LIB selects the build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
if os.environ.get("PAR_LIB"):
    par.library_path = lambda: os.environ["PAR_LIB"]
from pixel_art_remaster_gpu_b200 import synth
F, S, sub = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
base = synth.snes_stream(64, 256, 224)
frames = torch.from_numpy(np.concatenate([base] * (F // 64), 0)).cuda()
ctx = par.Remaster(0, 256, 224, F)
g = ctx.resolve_crossings(ctx.similarity_graph(frames))
for _ in range(4):
    ctx.raster(frames, g, S, sub)
torch.cuda.synchronize()

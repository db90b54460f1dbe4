import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
if os.environ.get("PAR_LIB"):
    par.library_path = lambda: os.environ["PAR_LIB"]
from pixel_art_remaster_gpu_b200 import synth
F, W, H = 1024, 256, 224
base = synth.snes_stream(64, W, H)
frames = torch.from_numpy(np.concatenate([base] * (F // 64), 0)).cuda()
ctx = par.Remaster(0, W, H, F)
g = ctx.resolve_crossings(ctx.similarity_graph(frames))
def best(fn, n=10, rounds=4):
    for _ in range(2): fn()
    torch.cuda.synchronize(); b = 1e9
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1) / n)
    return round(b, 4)
print(json.dumps({"lib": os.environ.get("PAR_LIB", "default"), "K3_1024_ms": best(lambda: ctx.cc_labels(g))}))

"""Scratch timing of each stage on a batch of G1 frames (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
if os.environ.get("PAR_LIB"):  # A/B runs: another build of the library
    par.library_path = lambda: os.environ["PAR_LIB"]
from pixel_art_remaster_gpu_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W, H = (256, 224) if len(sys.argv) <= 3 else (int(sys.argv[3]), int(sys.argv[4]))
uniq = min(F, 64)
base = synth.snes_stream(uniq, W, H)
frames = torch.from_numpy(np.concatenate([base] * (F // uniq), 0)).cuda()
F = frames.shape[0]
ctx = par.Remaster(0, W, H, F)
out = ctx._alloc(F, H, W, S, ("rgba", "graph", "graph_aux", "labels"))

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

px = F * W * H
for sub in (True, False):
    t = timeit(lambda: ctx.remaster(frames, S, sub, out={"rgba": out["rgba"], "graph": out["graph"], "graph_aux": out["graph_aux"]}))
    print(f"full sub={sub}: {t:.3f} ms for {F} frames -> {F / t * 1e3:,.0f} fps, {(3 + 1 + 4 * S * S) * px / t / 1e6:,.1f} GB/s algorithmic")
t = timeit(lambda: ctx.similarity_graph(frames)); print(f"K1 graph      : {t:.3f} ms  {4 * px / t / 1e6:,.1f} GB/s")
aux = ctx.similarity_graph(frames)
t = timeit(lambda: ctx.resolve_crossings(aux)); print(f"K2 crossings  : {t:.3f} ms  {2 * px / t / 1e6:,.1f} GB/s")
g = ctx.resolve_crossings(aux)
t = timeit(lambda: ctx.cc_labels(g)); print(f"K3 labels     : {t:.3f} ms  {5 * px / t / 1e6:,.1f} GB/s")
for sub in (True, False):
    t = timeit(lambda: ctx.raster(frames, g, S, sub)); print(f"K4 raster sub={sub}: {t:.3f} ms  {(4 + 4 * S * S) * px / t / 1e6:,.1f} GB/s")
print("smooth", ctx.smooth_stats())
ctx.no_tables = True
t = timeit(lambda: ctx.raster(frames, g, S, True)); print(f"K4 raster sub=True NO TABLES: {t:.3f} ms  {(4 + 4 * S * S) * px / t / 1e6:,.1f} GB/s")
ctx.no_tables = False
t = timeit(lambda: out["rgba"].fill_(7)); print(f"memset rgba   : {t:.3f} ms  {4 * S * S * px / t / 1e6:,.1f} GB/s")

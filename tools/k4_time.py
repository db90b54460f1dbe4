"""Scratch timing of the raster kernel alone (development aid): python tools/k4_time.py [frames] [scale]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
if os.environ.get("PAR_LIB"):  # A/B runs: another build of the library
    par.library_path = lambda: os.environ["PAR_LIB"]
from pixel_art_remaster_gpu_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W, H = 256, 224
base = synth.snes_stream(64, W, H)
if os.environ.get("K4_SPARSE"):  # flat 4x4 blocks: few smoothed cells per warp (the compacted pass of the raster kernel)
    base = np.ascontiguousarray(np.repeat(np.repeat(base[:, ::4, ::4], 4, 1), 4, 2))
if os.environ.get("K4_NOISE"):  # 3-colour noise: every key, many blends whose vertex the neighbour does not have (geometric path)
    rng = np.random.default_rng(5)
    base = np.ascontiguousarray(rng.integers(0, 256, (3, 3), dtype=np.uint8)[rng.integers(0, 3, (64, H, W))])
frames = torch.from_numpy(np.concatenate([base] * (F // 64), 0)).cuda()
ctx = par.Remaster(0, W, H, F)
g = ctx.resolve_crossings(ctx.similarity_graph(frames))

def timeit(fn, n=20, rounds=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(rounds):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    return best

px = F * W * H
for sub in (True, False):
    t = timeit(lambda: ctx.raster(frames, g, S, sub))
    print(f"K4 raster s={S} sub={sub}: {t:.3f} ms / {F} frames  {(4 + 4 * S * S) * px / t / 1e6:,.1f} GB/s  ({100 * (4 + 4 * S * S) * px / t / 1e6 / 6545.6:.1f} % of 6545.6)")
print("smooth", ctx.smooth_stats())

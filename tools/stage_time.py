"""Per-stage device times of the path on the bench stream (development aid; PAR_LIB=<variant .so> for A/B runs):
python tools/stage_time.py [frames]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pixel_art_remaster_gpu_b200 as par
if os.environ.get("PAR_LIB"):
    par.library_path = lambda: os.environ["PAR_LIB"]
from pixel_art_remaster_gpu_b200 import synth

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
W, H, S = 256, 224, 4
frames = torch.from_numpy(synth.snes_stream(F, W, H)).cuda()
ctx = par.Remaster(0, W, H, F)


def best(fn, n=10, rounds=4):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    b = 1e9
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1) / n)
    return round(b, 4)


aux = ctx.similarity_graph(frames)
g = ctx.resolve_crossings(aux)
row = {"lib": os.environ.get("PAR_LIB", "default"), "frames": F}
row["K1_ms"] = best(lambda: ctx.similarity_graph(frames))
row["K2_ms"] = best(lambda: ctx.resolve_crossings(aux))
row["K3_1024_ms"] = best(lambda: ctx.cc_labels(g[:1024]))
row["K4_ms"] = best(lambda: ctx.raster(frames, g, S, True))
row["K4_off_ms"] = best(lambda: ctx.raster(frames, g, S, False))
out = {"rgba": torch.empty((F, S * H, S * W, 4), dtype=torch.uint8, device="cuda"), "graph": torch.empty((F, H, W), dtype=torch.uint8, device="cuda")}
row["path_ms"] = best(lambda: ctx.remaster(frames, S, True, out=out))
print(json.dumps(row), flush=True)

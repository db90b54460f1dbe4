"""Per-stage device times INSIDE remaster() (par_profile_*: CUDA events around every stage launch), PAR_LIB=<variant .so> for A/B runs.  python tools/pipeline_stage_time.py [frames]"""
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
out = {"rgba": torch.empty((F, S * H, S * W, 4), dtype=torch.uint8, device="cuda"), "graph": torch.empty((F, H, W), dtype=torch.uint8, device="cuda")}
for _ in range(3):
    ctx.remaster(frames, S, True, out=out)
torch.cuda.synchronize()
ctx.profile(True)
n = 10
for _ in range(n):
    ctx.remaster(frames, S, True, out=out)
prof = ctx.profile_read()
ms = [prof[k][0] for k in ctx.STAGES]
ctx.profile(False)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(n):
    ctx.remaster(frames, S, True, out=out)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.environ.get("PAR_LIB", "default"), "frames": F,
                  "K1_ms": round(ms[0] / n, 4), "K2_ms": round(ms[1] / n, 4), "K4_ms": round(ms[4] / n, 4), "path_ms": round(e0.elapsed_time(e1) / n, 4)}))

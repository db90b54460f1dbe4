"""Round-2 probe (GPU box): sub-batch sweep, per-stage times, output formats, end-to-end per format, pinned-copy peak.
Prints one JSON object per measurement; not a bench value (bench.py is)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth

W, H, S = 256, 224, 4
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
host = torch.from_numpy(synth.snes_stream(N, W, H))
frames = host.to(dev)
ctx = par.Remaster(0, W, H, N)


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {"rgba": torch.empty((N, S * H, S * W, 4), dtype=torch.uint8, device=dev), "graph": torch.empty((N, H, W), dtype=torch.uint8, device=dev)}
for sub in (0, 32, 64, 128, 256, 512, 1024):
    ctx.set_sub_batch(sub)
    ms = timed(lambda: ctx.remaster(frames, S, True, out=out))
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(3):
        ctx.remaster(frames, S, True, out=out)
    prof = {k: round(v[0] / 3, 4) for k, v in ctx.profile_read().items() if v[1]}
    ctx.profile(False)
    print(json.dumps({"sub_batch": sub, "ms_per_step": round(ms, 4), "fps": round(N / ms * 1e3), "stage_ms": prof}), flush=True)
ctx.set_sub_batch(0)
del out
for fmt, name in ((par.OUT_RGBA8, "rgba8"), (par.OUT_BGR8, "bgr8"), (par.OUT_INDEX8, "index8")):
    o = ctx._alloc(N, H, W, S, ("rgba", "graph"), fmt)
    ms = timed(lambda: ctx.remaster(frames, S, True, out=o, out_format=fmt))
    ctx.profile(True)
    ctx.profile_read()
    for _ in range(3):
        ctx.remaster(frames, S, True, out=o, out_format=fmt)
    prof = {k: round(v[0] / 3, 4) for k, v in ctx.profile_read().items() if v[1]}
    ctx.profile(False)
    print(json.dumps({"format": name, "ms_per_step": round(ms, 4), "fps": round(N / ms * 1e3), "stage_ms": prof}), flush=True)
    del o
# end to end per format, pinned host buffers
pin_in = host.clone().pin_memory()
for fmt, name in ((par.OUT_RGBA8, "rgba8"), (par.OUT_BGR8, "bgr8"), (par.OUT_INDEX8, "index8")):
    o = ctx._alloc(N, H, W, S, ("rgba",), fmt, host=True)
    ctx.remaster_host(pin_in, S, True, out=o, out_format=fmt)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ctx.remaster_host(pin_in, S, True, out=o, out_format=fmt)
    dt = (time.perf_counter() - t0) / reps
    d2h = sum(v.numel() * v.element_size() for v in o.values())
    print(json.dumps({"e2e_format": name, "fps": round(N / dt), "ms": round(dt * 1e3, 2), "d2h_gbs": round(d2h / dt / 1e9, 2), "d2h_bytes": d2h}), flush=True)
    del o
# pinned copy peaks
buf_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
buf_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for name, fn in (("d2h", lambda: buf_h.copy_(buf_d, non_blocking=True)), ("h2d", lambda: buf_d.copy_(buf_h, non_blocking=True))):
    ms = timed(fn, reps=5, warm=1)
    print(json.dumps({"pinned_copy": name, "gbs": round((1 << 30) / ms / 1e6, 2)}), flush=True)
ctx.close()

"""Config C4 (BASELINE.json): one 4096x4096 synthetic pixel-art map, scale 4, labels on, tiled into horizontal strips
over the visible GPUs (par_group: strips + 40-row aprons by peer copies, labels stitched exactly).  Host image in, host
results out (pinned buffers), wall clock of par_group_remaster_host.  python tools/bench_c4.py [n_gpus] [reps]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W = H = 4096
img = torch.from_numpy(synth.pixel_art_map(W, H, synth.BASE_SEED + 4)).pin_memory().numpy()
pinned = {"rgba": torch.empty((4 * H, 4 * W, 4), dtype=torch.uint8).pin_memory().numpy(), "graph": torch.empty((H, W), dtype=torch.uint8).pin_memory().numpy(),
          "labels": torch.empty((H, W), dtype=torch.int32).pin_memory().numpy()}
res = {}
for strips, devices in ((1, [0]), (n, list(range(n)))):
    if strips in res:
        continue
    with par.RemasterGroup(devices, W, H, 4) as g:
        g.remaster_host(img, out=pinned)  # warm-up: tables, staging buffers
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = g.remaster_host(img, out=pinned)
            t.append(time.perf_counter() - t0)
    res[strips] = min(t)
    print("%d strip(s) on %d GPU(s): %.1f ms per 4096x4096 map (%.1f Mpx/s source, %.2f GB/s of results to the host)"
          % (strips, len(set(devices)), 1e3 * min(t), W * H / min(t) / 1e6, (W * H * (64 + 1 + 4)) / min(t) / 1e9))
print(json.dumps({"config": "C4 4096x4096 s=4 labels on, host in / host out", "seconds": {str(k): v for k, v in res.items()}}))

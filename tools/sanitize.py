"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, TMA and plain
tile paths, odd sizes, labels, polygons, launch_kernel, two strips."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pixel_art_remaster_gpu_b200 as par
from pixel_art_remaster_gpu_b200 import synth
for (W, H, S) in ((96, 80, 4), (50, 37, 8), (65, 33, 3)):
    frames = torch.from_numpy(synth.snes_stream(2, W, H, first_seed=7)).cuda()
    with par.Remaster(0, W, H, 2) as ctx:
        for no_tma in (False, True):
            out = ctx.remaster(frames, S, True, want=("rgba", "graph", "graph_aux", "labels", "polygons"), no_tma=no_tma)
            out = ctx.remaster(frames, S, True, want=("rgba",), no_tma=no_tma)  # smoothing tables
            ctx.no_tables = True
            out = ctx.remaster(frames, S, True, want=("rgba",), no_tma=no_tma)  # geometric path for every smoothed cell
            ctx.no_tables = False
            out = ctx.remaster(frames, S, False, want=("rgba",), no_tma=no_tma)  # hull cells only
            for fmt in (par.OUT_BGR8, par.OUT_INDEX8):  # output formats (+ the palette kernels)
                out = ctx.remaster(frames, S, True, no_tma=no_tma, out_format=fmt)
        lab = ctx.remaster(frames, 1, False, want=("graph", "labels"))
        wl, wb, nodes, total = ctx.border_walks(lab["graph"], lab["labels"])
        ctx.walk_splines(wl, wb, nodes, total, 4)
        torch.cuda.synchronize()
# 3-colour noise: every key, cells with three and four link descriptors (second table pass), blends whose vertex the
# neighbour does not have (geometric path, and from the second pass the exact tile resolve)
rng = np.random.default_rng(5)
noise = torch.from_numpy(rng.integers(0, 256, (3, 3), dtype=np.uint8)[rng.integers(0, 3, (2, 70, 90))]).cuda()
with par.Remaster(0, 90, 70, 2) as ctx:
    for S in (4, 6):
        out = ctx.remaster(noise, S, True, want=("rgba",))
    torch.cuda.synchronize()
img = synth.adversarial_sprite(96, 100, 3)
par.launch_kernel(img, True)
with par.RemasterGroup([0, 0, 0], 96, 130, 4) as g:
    big = synth.adversarial_sprite(96, 130, 3)
    g.remaster_host(big, want=("rgba", "graph", "labels"))
    for st in g.strips():  # the device-resident entry: halo rows by peer copies, on-device label stitch
        lo, hi = st["own"]
        st["bgr"][lo - st["load"][0]:hi - st["load"][0]].copy_(torch.from_numpy(big[lo:hi]))
    torch.cuda.synchronize()
    g.remaster_device(True)
print("sanitize run done")

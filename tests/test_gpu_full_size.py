"""BASELINE.json's full-size configurations, checked through size-independent properties (the oracle
is only run on windows it finishes in seconds):
  C3  stream of 4096 256x224 frames in ONE batch == the same frames remastered one at a time on a fresh
      context with the smoothing tables off; a sample of them == the oracle
  C4  one 4096x4096 map: 8 strips with aprons and stitched labels == the whole image on one context;
      random windows (with the exact 40-row/column apron) == the oracle
"""
import numpy as np
import pytest

from pixel_art_remaster_gpu_b200 import synth

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def test_c3_stream_of_4096_frames(lib, oracle):
    torch = _torch()
    n, W, H = 4096, 256, 224
    base = synth.snes_stream(256, W, H, first_seed=synth.BASE_SEED)           # 256 distinct frames ...
    frames_np = np.concatenate([base[(np.arange(256) * 7 + k) % 256] for k in range(n // 256)], 0)  # ... shuffled into 4096
    frames = torch.from_numpy(frames_np).cuda()
    with lib.Remaster(0, W, H, n) as ctx:
        out = ctx.remaster(frames, scale=4, subdivide=True, want=("rgba", "graph"))
        torch.cuda.synchronize()
        # linearity over the batch axis: any sub-batch gives the same frames
        sub = ctx.remaster(frames[1000:1016], scale=4, subdivide=True, want=("rgba", "graph"))
        assert torch.equal(sub["rgba"], out["rgba"][1000:1016]) and torch.equal(sub["graph"], out["graph"][1000:1016])
        # identical inputs give identical outputs wherever they sit in the batch (frames repeat every 256 modulo the shuffle)
        a, b = 5, 256 + (5 - 7 * 0) % 256  # frame k of block 0 is base[(7k) % 256]; find its twin in block 1
        src0 = (7 * a + 0) % 256
        twin = [k for k in range(256) if (7 * k + 1) % 256 == src0][0] + 256
        assert torch.equal(out["rgba"][a], out["rgba"][twin])
    with lib.Remaster(0, W, H, 1) as fresh:                                     # geometric path, one frame at a time
        fresh.no_tables = True
        for k in (0, 777, 2048, 4095):
            one = fresh.remaster(frames[k:k + 1], scale=4, subdivide=True, want=("rgba", "graph"))
            assert torch.equal(one["rgba"][0], out["rgba"][k]) and torch.equal(one["graph"][0], out["graph"][k])
    for k in (0, 4095):
        want = oracle.pipeline(frames_np[k], scale=4, want=("graph", "raster"))
        assert np.array_equal(out["graph"][k].cpu().numpy(), want["graph"])
        assert np.array_equal(out["rgba"][k].cpu().numpy(), want["raster"])


def test_c4_map_4096_tiled_equals_whole(lib, oracle):
    torch = _torch()
    W = H = 4096
    img = synth.pixel_art_map(W, H, synth.BASE_SEED + 4)
    n_gpus = torch.cuda.device_count()
    with lib.Remaster(0, W, H, 1) as ctx:
        whole = ctx.remaster(torch.from_numpy(img[None]).cuda(), scale=4, subdivide=True, want=("rgba", "graph", "labels"))
        w_graph, w_lab = whole["graph"][0].cpu().numpy(), whole["labels"][0].cpu().numpy()
        w_rgba = whole["rgba"][0].cpu().numpy()
        del whole
    torch.cuda.empty_cache()
    with lib.RemasterGroup([k % n_gpus for k in range(8)], W, H, 4) as grp:
        tiled = grp.remaster_host(img, subdivide=True, want=("rgba", "graph", "labels"))
    assert np.array_equal(tiled["graph"], w_graph)
    assert np.array_equal(tiled["labels"], w_lab)          # stitched labels == single-image labels, bit for bit
    assert np.array_equal(tiled["rgba"], w_rgba)
    # labels are canonical: every label is the index of a pixel that carries that label and is the smallest such
    flat = w_lab.reshape(-1)
    assert (flat <= np.arange(flat.size)).all() and (flat[flat] == flat).all()
    # windows against the oracle: the exact dependency radius is 37, so a 40-pixel apron makes the interior exact
    rng = np.random.default_rng(4)
    for _ in range(3):
        x0, y0 = int(rng.integers(40, W - 240)), int(rng.integers(40, H - 240))
        win = np.ascontiguousarray(img[y0 - 40:y0 + 200 + 40, x0 - 40:x0 + 200 + 40])
        want = oracle.pipeline(win, scale=4, want=("graph", "raster"))
        assert np.array_equal(want["graph"][40:-40, 40:-40], w_graph[y0:y0 + 200, x0:x0 + 200])
        assert np.array_equal(want["raster"][160:-160, 160:-160], w_rgba[4 * y0:4 * (y0 + 200), 4 * x0:4 * (x0 + 200)])


def test_more_frames_than_one_launch_can_index(lib, oracle):
    """A batch of more than 65 535 frames (the grid.z limit: every stage splits its launches): 70 000 frames of 8 x 8,
    whole path with labels, all three output formats — equal to the same frames run in two halves, a sample against the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    F, W, H, S = 70_000, 8, 8, 2
    rng = np.random.default_rng(9)
    pal = rng.integers(0, 256, (5, 3), dtype=np.uint8)
    frames_np = pal[rng.integers(0, 5, (F, H, W))]
    frames = torch.from_numpy(frames_np).cuda()
    with lib.Remaster(0, W, H, F) as c:
        whole = c.remaster(frames, scale=S, subdivide=True, want=("rgba", "graph", "graph_aux", "labels"))
        idx = c.remaster(frames, scale=S, subdivide=True, out_format=lib.OUT_INDEX8)
        bgr = c.remaster(frames, scale=S, subdivide=True, out_format=lib.OUT_BGR8)["rgba"]
        half = F // 2
        for lo, hi in ((0, half), (half, F)):
            part = c.remaster(frames[lo:hi], scale=S, subdivide=True, want=("rgba", "graph", "graph_aux", "labels"))
            for k in part:
                assert torch.equal(part[k], whole[k][lo:hi]), (k, lo)
        assert torch.equal(bgr, whole["rgba"][..., [2, 1, 0]])
        for lo in range(0, F, 10_000):  # (expanded in pieces: the gather index is 8 bytes per pixel)
            hi = min(lo + 10_000, F)
            assert np.array_equal(lib.Remaster.expand_indexed(idx["rgba"][lo:hi], idx["palette"][lo:hi]), whole["rgba"][lo:hi].cpu().numpy())
        for k in (0, 65_534, 65_535, 65_536, F - 1):
            want = oracle.pipeline(frames_np[k], scale=S, want=("graph_aux", "graph", "labels", "raster"))
            assert np.array_equal(whole["graph"][k].cpu().numpy(), want["graph"]) and np.array_equal(whole["labels"][k].cpu().numpy(), want["labels"])
            assert np.array_equal(whole["rgba"][k].cpu().numpy(), want["raster"]), k

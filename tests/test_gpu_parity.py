"""Parity of the CUDA path (through the C ABI) against the CPU oracle, stage by stage.

Bars (BASELINE.json north_star): similarity-graph edges, crossing decisions and CC labels BIT-EXACT;
polygon vertices equal (exact dyadic arithmetic; tolerance written below is 1e-5 relative as stated,
the assertion used is the stronger equality); raster within +-1 LSB per channel (asserted: equal).
"""
import numpy as np
import pytest

from conftest import valid_vertex_mask
from pixel_art_remaster_gpu_b200 import synth

pytestmark = pytest.mark.gpu

VERTEX_RTOL = 1e-5
RASTER_LSB = 1


def _dev(ctx, frames):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(frames))
    return t.to(ctx.device)


def _cases():
    return [
        ("g1_small", synth.snes_frame(96, 80, synth.BASE_SEED + 1)),
        ("g1_c1", synth.snes_frame(256, 224, synth.BASE_SEED)),
        ("g5_small", synth.adversarial_sprite(192, 160)),
        ("odd_size", synth.snes_frame(50, 37, 5)),
        ("one_tile_minus", synth.snes_frame(63, 31, 6)),
        ("one_tile_plus", synth.snes_frame(65, 33, 7)),
        ("tiny", synth.snes_frame(3, 2, 8)),
        ("single_pixel", synth.snes_frame(1, 1, 9)),
        ("single_row", synth.snes_frame(40, 1, 10)),
        ("single_column", synth.snes_frame(1, 40, 11)),
    ]


@pytest.mark.parametrize("name,img", _cases(), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("no_tma", [False, True], ids=["tma", "plain"])
def test_graph_stages_bit_exact(ctx, oracle, name, img, no_tma):
    want = oracle.pipeline(img, want=("graph_aux", "graph"))
    frames = _dev(ctx, img[None])
    aux = ctx.similarity_graph(frames, no_tma=no_tma)
    g = ctx.resolve_crossings(aux, no_tma=no_tma)
    assert np.array_equal(aux[0].cpu().numpy(), want["graph_aux"]), "stage A+B differs"
    assert np.array_equal(g[0].cpu().numpy(), want["graph"]), "stage C differs"


@pytest.mark.parametrize("name,img", _cases(), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("no_tma", [False, True], ids=["tma", "plain"])
def test_whole_path_graphs_bit_exact(ctx, oracle, name, img, no_tma):
    """Both graphs as the whole path (par_remaster_device) leaves them — same bytes as the stage entry points and as the oracle."""
    want = oracle.pipeline(img, want=("graph_aux", "graph"))
    out = ctx.remaster(_dev(ctx, img[None]), scale=1, subdivide=False, want=("graph", "graph_aux"), no_tma=no_tma)
    assert np.array_equal(out["graph_aux"][0].cpu().numpy(), want["graph_aux"]), "stage A+B differs"
    assert np.array_equal(out["graph"][0].cpu().numpy(), want["graph"]), "stage C differs"


def test_whole_path_graphs_on_hard_inputs(lib, oracle):
    """The graph stages on inputs chosen against them: a checkerboard (every block ambiguous and falling through to the
    curve-length walks), 2- and 4-colour noise (all rules, chains that cross tile seams), near-black pixels on the borders,
    widths that are not multiples of 4 or of the tile, batches — whole path and stage entries against the oracle, frame by frame."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    rng = np.random.default_rng(11)
    for W, H, F in ((64, 32, 3), (65, 33, 3), (130, 70, 4), (63, 95, 2), (200, 129, 3), (5, 300, 2), (301, 6, 2)):
        frames = np.zeros((F, H, W, 3), np.uint8)
        yy, xx = np.mgrid[0:H, 0:W]
        frames[0][(yy + xx) % 2 == 0] = (250, 250, 250)                                      # checkerboard
        pal = np.array([[0, 0, 0], [3, 2, 4], [200, 40, 90], [250, 250, 250]], np.uint8)
        frames[1] = pal[rng.integers(0, 2, (H, W)) * 2]                                       # 2-colour noise (black / colour)
        for k in range(2, F):
            frames[k] = pal[rng.integers(0, 4, (H, W))]                                       # near-black + colours
        for no_tma in (False, True):
            with lib.Remaster(0, W, H, F) as c:
                out = c.remaster(torch.from_numpy(frames).cuda(), scale=1, subdivide=False, want=("graph", "graph_aux"), no_tma=no_tma)
                aux = c.similarity_graph(torch.from_numpy(frames).cuda(), no_tma=no_tma)
                assert torch.equal(aux, out["graph_aux"])
                assert torch.equal(c.resolve_crossings(aux, no_tma=no_tma), out["graph"])
            for k in range(F):
                want = oracle.pipeline(frames[k], want=("graph_aux", "graph"))
                assert np.array_equal(out["graph_aux"][k].cpu().numpy(), want["graph_aux"]), (W, H, k, no_tma)
                assert np.array_equal(out["graph"][k].cpu().numpy(), want["graph"]), (W, H, k, no_tma)


@pytest.mark.parametrize("w,h", [(1, 1), (2, 3), (4, 4), (7, 5), (63, 31), (64, 32), (65, 33), (128, 64), (129, 35), (200, 70)])
@pytest.mark.parametrize("no_tma", [False, True], ids=["tma", "plain"])
def test_graph_black_and_dark_pixels_on_the_border(ctx, oracle, w, h, no_tma):
    """The graph kernel stages pixels outside the image as zeros (black) and clears the links that leave the image when
    the bytes are assembled: black and near-black pixels (similar to black: Y <= 5, U <= 7, V <= 6) on the image border,
    in the corners and at tile edges must not link outwards, and blocks across the border must not lose diagonals."""
    rng = np.random.default_rng(1000 * w + h)
    palette = np.array([[0, 0, 0], [1, 2, 1], [4, 4, 4], [3, 0, 5], [200, 30, 90], [0, 0, 0], [2, 2, 2], [250, 250, 250]], np.uint8)
    for k in range(3):
        img = palette[rng.integers(0, len(palette) if k else 4, (h, w))]
        if k == 2:
            img[:] = 0  # an all-black frame: every in-image link set, none outwards
        img = np.ascontiguousarray(img)
        want = oracle.pipeline(img, want=("graph_aux", "graph"))
        aux = ctx.similarity_graph(_dev(ctx, img[None]), no_tma=no_tma)
        g = ctx.resolve_crossings(aux, no_tma=no_tma)
        assert np.array_equal(aux[0].cpu().numpy(), want["graph_aux"]), "stage A+B differs"
        assert np.array_equal(g[0].cpu().numpy(), want["graph"]), "stage C differs"


@pytest.mark.parametrize("name,img", _cases(), ids=[c[0] for c in _cases()])
def test_cc_labels_bit_exact(ctx, oracle, name, img):
    want = oracle.pipeline(img, want=("graph", "labels"))
    import torch
    g = torch.from_numpy(want["graph"][None]).to(ctx.device)
    lab = ctx.cc_labels(g)
    assert np.array_equal(lab[0].cpu().numpy(), want["labels"])


def test_cc_labels_on_hard_graphs(lib, oracle):
    """The labeller on inputs chosen against union-find: one component that snakes through the whole frame, a checkerboard
    graph (diagonal links only), noise, widths that are not multiples of the tile, and a batch."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    rng = np.random.default_rng(21)
    for W, H in ((256, 256), (257, 256), (255, 257), (33, 70), (31, 3), (1, 50), (400, 160)):
        imgs = []
        snake = np.zeros((H, W, 3), np.uint8)
        snake[...] = (10, 200, 30)
        for x in range(1, W - 1, 4):
            snake[1:H - 1, x] = (250, 20, 20)
            if x + 4 < W - 1:
                snake[(H - 2) if (x // 4) % 2 else 1, x:x + 5] = (250, 20, 20)
        imgs.append(snake)
        chk = np.zeros((H, W, 3), np.uint8)
        yy, xx = np.mgrid[0:H, 0:W]
        chk[(yy + xx) % 2 == 0] = (255, 255, 255)
        imgs.append(chk)
        pal = rng.integers(0, 256, (3, 3), dtype=np.uint8)
        imgs.append(pal[rng.integers(0, 3, (H, W))])
        graphs = np.stack([oracle.pipeline(im, want=("graph",))["graph"] for im in imgs])
        want = np.stack([oracle.cc_labels(g) for g in graphs])
        with lib.Remaster(0, W, H, len(imgs)) as c:
            g = torch.from_numpy(graphs).cuda()
            assert np.array_equal(c.cc_labels(g).cpu().numpy(), want), (W, H)


@pytest.mark.parametrize("name,img", _cases(), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("subdivide", [False, True], ids=["hull", "subdivided"])
def test_polygons_equal(ctx, oracle, name, img, subdivide):
    want = oracle.pipeline(img, subdivide=subdivide, want=("graph", "poly", "poly_count"))
    import torch
    frames = _dev(ctx, img[None])
    g = torch.from_numpy(want["graph"][None]).to(ctx.device)
    poly, cnt = ctx.polygons(frames, g, subdivide=subdivide)
    poly, cnt = poly[0].cpu().numpy(), cnt[0].cpu().numpy()
    assert np.array_equal(cnt, want["poly_count"])
    m = valid_vertex_mask(cnt)
    np.testing.assert_allclose(poly[m], want["poly"][m], rtol=VERTEX_RTOL, atol=0)
    assert np.array_equal(poly[m], want["poly"][m])  # in fact exact


@pytest.mark.parametrize("name,img", _cases(), ids=[c[0] for c in _cases()])
@pytest.mark.parametrize("scale,subdivide", [(4, True), (4, False), (8, True), (1, True), (2, True), (3, True), (6, False), (5, True), (7, True),
                                             (5, False), (7, False), (6, True), (8, False)])
def test_raster_within_one_lsb(ctx, oracle, name, img, scale, subdivide):
    want = oracle.pipeline(img, subdivide=subdivide, scale=scale, want=("graph", "raster"))
    import torch
    frames = _dev(ctx, img[None])
    g = torch.from_numpy(want["graph"][None]).to(ctx.device)
    for no_tma in (False, True):
        rgba = ctx.raster(frames, g, scale=scale, subdivide=subdivide, no_tma=no_tma)[0].cpu().numpy()
        diff = np.abs(rgba.astype(np.int16) - want["raster"].astype(np.int16))
        assert diff.max() <= RASTER_LSB, "%d output pixels differ by more than 1 LSB" % int((diff.max(-1) > RASTER_LSB).sum())
        assert diff.max() == 0


@pytest.mark.parametrize("scale", [4, 8, 3, 5, 7])
def test_raster_exact_slow_path(ctx, oracle, scale):
    """Cells that reach beyond their sample mask are rasterized by an exact slow path; force every cell
    through it (PAR_FLAG_DEBUG_WIDE) and require the same image."""
    img = synth.snes_frame(96, 80, synth.BASE_SEED + 3)
    want = oracle.pipeline(img, scale=scale, want=("graph", "raster"))
    import torch
    g = torch.from_numpy(want["graph"][None]).to(ctx.device)
    rgba = ctx.raster(_dev(ctx, img[None]), g, scale=scale, subdivide=True, debug_wide=True)[0].cpu().numpy()
    assert np.array_equal(rgba, want["raster"])


def test_padded_rows_and_flip(ctx, oracle):
    """widthstep > 3*width (OpenCV row alignment) and the top-scanline-first output option."""
    base = synth.snes_frame(50, 37, 21)
    img = synth.pad_rows(base, 152)
    want = oracle.pipeline(img, scale=4, want=("graph_aux", "graph", "raster", "poly", "poly_count"))
    import torch
    buf = torch.from_numpy(np.lib.stride_tricks.as_strided(img, shape=(37, 152), strides=(152, 1)).copy()).to(ctx.device)
    frames = torch.as_strided(buf, (1, 37, 50, 3), (37 * 152, 152, 3, 1))
    out = ctx.remaster(frames, scale=4, subdivide=True, want=("rgba", "graph", "graph_aux", "polygons"))
    assert np.array_equal(out["graph_aux"][0].cpu().numpy(), want["graph_aux"])
    assert np.array_equal(out["graph"][0].cpu().numpy(), want["graph"])
    assert np.array_equal(out["rgba"][0].cpu().numpy(), want["raster"])
    cnt = out["poly_count"][0].cpu().numpy()
    m = valid_vertex_mask(cnt)
    assert np.array_equal(cnt, want["poly_count"]) and np.array_equal(out["polygons"][0].cpu().numpy()[m], want["poly"][m])
    flipped = ctx.remaster(frames, scale=4, subdivide=True, want=("rgba",), flip_output=True)["rgba"][0].cpu().numpy()
    assert np.array_equal(flipped, want["raster"][::-1])


def test_batch_of_frames_and_host_path(ctx, oracle):
    """A batch is processed frame-independently (blockIdx.z = frame); host-buffer entry point agrees."""
    frames_np = synth.snes_stream(5, 96, 80, first_seed=100)
    frames = _dev(ctx, frames_np)
    out = ctx.remaster(frames, scale=4, subdivide=True, want=("rgba", "graph", "labels"))
    import torch
    host = ctx.remaster_host(torch.from_numpy(frames_np).pin_memory(), scale=4, subdivide=True, want=("rgba", "graph", "labels"))
    for k in range(5):
        want = oracle.pipeline(frames_np[k], scale=4, want=("graph", "labels", "raster"))
        assert np.array_equal(out["graph"][k].cpu().numpy(), want["graph"])
        assert np.array_equal(out["labels"][k].cpu().numpy(), want["labels"])
        assert np.array_equal(out["rgba"][k].cpu().numpy(), want["raster"])
        assert np.array_equal(host["rgba"][k].numpy(), want["raster"])
        assert np.array_equal(host["labels"][k].numpy(), want["labels"])


def test_config2_full_pipeline_320x240_s8(ctx, oracle):
    """BASELINE config 2: 320x240, CC labels, subdivision, raster at 8x."""
    img = synth.snes_frame(320, 240, synth.BASE_SEED + 2)
    want = oracle.pipeline(img, scale=8, want=("graph", "labels", "raster"))
    out = ctx.remaster(_dev(ctx, img[None]), scale=8, subdivide=True, want=("rgba", "graph", "labels"))
    assert np.array_equal(out["graph"][0].cpu().numpy(), want["graph"])
    assert np.array_equal(out["labels"][0].cpu().numpy(), want["labels"])
    assert np.array_equal(out["rgba"][0].cpu().numpy(), want["raster"])


def test_config5_adversarial_512x448(ctx, oracle):
    """BASELINE config 5: dithered checkerboard / dense diagonals / islands / long thin components."""
    img = synth.adversarial_sprite(512, 448)
    want = oracle.pipeline(img, scale=4, want=("graph_aux", "graph", "labels", "raster"))
    n_amb = oracle.resolve_crossings(want["graph_aux"])[1]
    assert n_amb > 10000  # the input really is adversarial
    out = ctx.remaster(_dev(ctx, img[None]), scale=4, subdivide=True, want=("rgba", "graph", "graph_aux", "labels"))
    assert np.array_equal(out["graph_aux"][0].cpu().numpy(), want["graph_aux"])
    assert np.array_equal(out["graph"][0].cpu().numpy(), want["graph"])
    assert np.array_equal(out["labels"][0].cpu().numpy(), want["labels"])
    assert np.array_equal(out["rgba"][0].cpu().numpy(), want["raster"])


def test_exhaustive_yuv_words(ctx, oracle):
    """All 2^24 colours through the graph kernel's conversion: a 4096x4096 frame holding every colour once;
    neighbours differ by one step in byte 0 / byte 1 so the edges exercise the thresholds.  The WHOLE device graph
    (every pixel, so every colour's packed word takes part in up to eight comparisons) is compared with the graph
    numpy builds from the oracle's conversion table for all 2^24 colours (tests/np_graph.py, itself pinned against the
    oracle's stages A+B), and a slab of it with the oracle's own stages A+B."""
    from np_graph import all_colours_frame, graph_aux_from_yuv
    img = all_colours_frame()
    want = graph_aux_from_yuv(oracle.yuv_all(True).reshape(4096, 4096))
    slab = np.ascontiguousarray(img[1024:1024 + 256])
    assert np.array_equal(oracle.trivial_crossings(oracle.similarity_graph(slab))[1:-1], want[1025:1024 + 255])
    big = ctx._torch.from_numpy(img[None]).to(ctx.device)
    for no_tma in (False, True):
        aux = ctx.similarity_graph(big, no_tma=no_tma)[0].cpu().numpy()
        bad = np.argwhere(aux != want)
        assert len(bad) == 0, (no_tma, len(bad), bad[:5])


def test_smoothing_tables_equal_the_geometric_path(lib, oracle):
    """Stage E through the precomputed smoothing tables (XOR of CUT and LINK pieces) gives exactly the image of
    the geometric path (PAR_FLAG_NO_SMOOTH_TABLES: polygon built and rasterized per cell) at every scale, on
    ordinary frames, on the adversarial sprite and on raw noise, and both equal the oracle."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    rng = np.random.default_rng(77)
    pal = rng.integers(0, 256, (3, 3), dtype=np.uint8)
    noise = pal[rng.integers(0, 3, (2, 96, 128))]                      # 3-colour noise: every key, not-found blends
    noise[1, :, :] = (pal[rng.integers(0, 2, (96, 128))])              # 2-colour noise
    sets = {"snes": synth.snes_stream(3, 128, 96, first_seed=500),
            "adversarial": synth.adversarial_sprite(128, 96, 9)[None],
            "noise": np.ascontiguousarray(noise)}
    for name, frames_np in sets.items():
        frames = torch.from_numpy(frames_np).cuda()
        for scale in (1, 2, 3, 4, 5, 6, 7, 8):
            with lib.Remaster(0, 128, 96, frames_np.shape[0]) as c:
                tab = c.remaster(frames, scale=scale, subdivide=True)["rgba"].cpu().numpy()
                st = c.smooth_stats()
                assert st["smoothed"] > 0
                if name == "snes":
                    assert st["geometric"] < 0.01 * st["smoothed"], st       # the tables cover (almost) everything
                c.no_tables = True
                geo = c.remaster(frames, scale=scale, subdivide=True)["rgba"].cpu().numpy()
                st2 = c.smooth_stats()
                assert st2["geometric"] - st["geometric"] == st2["smoothed"] - st["smoothed"]
            assert np.array_equal(tab, geo), (name, scale, int((tab != geo).any(-1).sum()))
            if scale in (4, 5, 7, 8):
                assert np.array_equal(tab[0], oracle.pipeline(frames_np[0], scale=scale, want=("raster",))["raster"]), (name, scale)


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [4, 8])
def test_large_batch_of_small_frames(lib, oracle, scale):
    """A large batch (grid.z = 2047 frames, cycling through 16 distinct ones): every frame equals the oracle's image of
    its source frame — with the smoothing tables, on the geometric path, on the exact tile resolve and without TMA."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    W, H, n, distinct = 96, 80, 2047, 16
    base = synth.snes_stream(distinct, W, H, first_seed=synth.BASE_SEED + 900)
    want = torch.from_numpy(np.stack([oracle.pipeline(b, scale=scale, want=("raster",))["raster"] for b in base])).cuda()
    frames = torch.from_numpy(np.concatenate([base] * (n // distinct + 1))[:n]).cuda()
    index = torch.arange(n, device="cuda") % distinct
    with lib.Remaster(0, W, H, n) as c:
        g = c.resolve_crossings(c.similarity_graph(frames))
        for mode in ("tables", "geometric", "exact", "plain_loads"):
            c.no_tables = mode == "geometric"
            rgba = c.raster(frames, g, scale, True, debug_wide=mode == "exact", no_tma=mode == "plain_loads")
            bad = (rgba != want[index]).flatten(1).any(1)
            assert not bool(bad.any()), (mode, scale, bad.nonzero().flatten()[:8].tolist())
            del rgba


@pytest.mark.gpu
@pytest.mark.parametrize("scale,aa", [(4, 2), (2, 2), (2, 4), (1, 4), (1, 2), (3, 2)])
def test_antialiased_output_is_the_mean_of_the_supersampled_image(lib, oracle, scale, aa):
    """PAR_FLAG_AA2 / AA4: every output pixel is the per-channel mean (halves up) of aa x aa ordered-grid samples, i.e. of
    the point-sampled image at aa x the scale — compared here with the oracle's raster at that scale, averaged in numpy.
    (The reference's GL_MULTISAMPLE pattern is driver-defined, SURVEY §8(f)-2: this pins OUR rule, bit-exactly.)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    for frames_np in (synth.snes_stream(2, 96, 80, first_seed=31), synth.adversarial_sprite(96, 80, 5)[None]):
        F, H, W = frames_np.shape[:3]
        with lib.Remaster(0, W, H, F) as c:
            c.aa = aa
            for subdivide in (True, False):
                got = c.remaster(torch.from_numpy(frames_np).cuda(), scale=scale, subdivide=subdivide)["rgba"].cpu().numpy()
                assert got.shape == (F, H * scale, W * scale, 4)
                for k in range(F):
                    big = oracle.pipeline(frames_np[k], subdivide, True, scale * aa, ("raster",))["raster"].astype(np.uint32)
                    want = (big.reshape(H * scale, aa, W * scale, aa, 4).sum(axis=(1, 3)) + aa * aa // 2) // (aa * aa)
                    assert np.array_equal(got[k], want.astype(np.uint8)), (scale, aa, subdivide, k)
            c.aa = 1
    with lib.Remaster(0, 32, 32, 1) as c:  # scale x samples must be a supported sampling scale
        c.aa = 4
        with pytest.raises(Exception):
            c.remaster(torch.zeros((1, 32, 32, 3), dtype=torch.uint8).cuda(), scale=4, subdivide=True)


@pytest.mark.gpu
def test_border_walks(lib, oracle):
    """par_border_walks: bit-exact against the oracle on whole batches of synthetic frames, and against
    the reference's own walker through the golden graphs it survived (tests/golden/border_walks_small.json)."""
    import json, os
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    for frames_np in (synth.snes_stream(3, 96, 80, first_seed=77), synth.adversarial_sprite(120, 90, 3)[None], synth.snes_stream(2, 33, 21, first_seed=5)):
        F, H, W = frames_np.shape[:3]
        with lib.Remaster(0, W, H, F) as c:
            out = c.remaster(torch.from_numpy(frames_np).cuda(), scale=1, subdivide=False, want=("graph", "labels"))
            wl, wb, nodes, total = c.border_walks(out["graph"], out["labels"])
            torch.cuda.synchronize()
            for k in range(F):
                g = out["graph"][k].cpu().numpy()
                want = oracle.border_walks(g, out["labels"][k].cpu().numpy())
                got = c.walks_as_dict(wl, wb, nodes, k)
                assert got == want, (W, H, k)
                assert int(total[k]) == sum(len(w) for w in want.values())
                assert list(got) == sorted(got) and [int(wb[k].reshape(-1)[s]) for s in got] == list(np.cumsum([0] + [len(w) for w in got.values()])[:-1])
            # a capacity that is too small is reported, not overrun
            wl2, wb2, nodes2, total2 = c.border_walks(out["graph"], out["labels"], capacity_per_frame=8)
            torch.cuda.synchronize()
            assert torch.equal(total2, total) and torch.equal(wl2, wl)
            # splines through the walks (par_walk_splines): closed uniform quadratic B-spline over the node centres, exact
            # dyadic arithmetic -> equal to the numpy restatement bit for bit, at every sample count
            for samples in (1, 2, 4, 8):
                pts = c.walk_splines(wl, wb, nodes, total, samples).cpu().numpy()
                for k in range(F):
                    for start, walk in c.walks_as_dict(wl, wb, nodes, k).items():
                        b = int(wb[k].reshape(-1)[start])
                        want_pts = oracle.walk_spline(walk, W, samples)
                        got_pts = pts[k, b * samples:(b + len(walk)) * samples]
                        assert np.array_equal(got_pts.astype(np.float64), want_pts), (W, H, k, start, samples)
                        # the curve of a closed polygon stays inside the polygon's bounding box and passes through the edge midpoints
                        P = np.stack([np.asarray(walk) % W + 0.5, np.asarray(walk) // W + 0.5], -1)
                        assert np.array_equal(got_pts[::samples].astype(np.float64), (np.roll(P, 1, 0) + P) / 2)
            with pytest.raises(lib.RemasterError):
                c.walk_splines(wl, wb, nodes, total, 3)
    here = os.path.dirname(os.path.abspath(__file__))
    cases = json.load(open(os.path.join(here, "golden", "border_walks_small.json")))
    for case in cases:
        g = np.array(case["graph"], np.uint8).reshape(case["H"], case["W"])
        lab = oracle.cc_labels(g)
        first = {}
        for w in case["walks"]:
            if lab.reshape(-1)[w[0]] == w[0] and w[0] not in first:
                first[w[0]] = w
        with lib.Remaster(0, case["W"], case["H"], 1) as c:
            gd = torch.from_numpy(g[None]).cuda()
            labd = c.cc_labels(gd)
            assert np.array_equal(labd[0].cpu().numpy(), lab)
            got = c.walks_as_dict(*c.border_walks(gd, labd)[:3])
        assert got == first, case["name"]


@pytest.mark.gpu
def test_antialiased_output_on_every_code_path(lib):
    """The averaging sits in the resolve step of every variant: TMA and plain tile staging, flipped rows, the exact
    out-of-line tile resolve (PAR_FLAG_DEBUG_WIDE), the geometric path — all give the same anti-aliased image."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    frames_np = synth.snes_stream(2, 65, 33, first_seed=12)          # odd sizes: partial tiles
    frames = torch.from_numpy(frames_np).cuda()
    for scale, aa in ((2, 2), (4, 2), (2, 4)):
        with lib.Remaster(0, 65, 33, 2) as c:
            c.aa = aa
            g = c.resolve_crossings(c.similarity_graph(frames))
            base = c.raster(frames, g, scale, True).cpu().numpy()
            assert np.array_equal(c.raster(frames, g, scale, True, no_tma=True).cpu().numpy(), base)
            assert np.array_equal(c.raster(frames, g, scale, True, debug_wide=True).cpu().numpy(), base)
            assert np.array_equal(c.raster(frames, g, scale, True, flip_output=True).cpu().numpy(), base[:, ::-1])
            c.no_tables = True
            assert np.array_equal(c.raster(frames, g, scale, True).cpu().numpy(), base)


@pytest.mark.gpu
def test_c_abi_error_behaviour(lib):
    """Bad jobs are refused with a status and a message, nothing is launched and the context stays usable."""
    import ctypes as C
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    L = lib.load_library()
    frames = torch.from_numpy(synth.snes_stream(2, 40, 30, first_seed=3)).cuda()
    with lib.Remaster(0, 40, 30, 2) as c:
        good = c.remaster(frames, scale=2, subdivide=True)["rgba"].clone()
        launches = c.launch_count
        rgba = torch.empty((2, 60, 80, 4), dtype=torch.uint8, device="cuda")

        def job(**kw):
            j = lib.ParJob()
            j.bgr, j.width, j.height, j.widthstep, j.frame_stride, j.n_frames, j.scale, j.flags = frames.data_ptr(), 40, 30, 120, 0, 2, 2, 1
            j.rgba = rgba.data_ptr()
            for k, v in kw.items():
                setattr(j, k, v)
            return j

        INVALID, CAPACITY = 1, 4
        for bad, status in ((job(n_frames=0), INVALID), (job(width=0), INVALID), (job(height=-3), INVALID), (job(bgr=None), INVALID),
                            (job(widthstep=100), INVALID), (job(frame_stride=100), INVALID), (job(scale=0), INVALID), (job(scale=9), INVALID),
                            (job(out_format=3), INVALID), (job(out_format=2, flags=1 | 32), INVALID),   # unknown format; indexed + anti-aliased
                            (job(rgba=rgba.data_ptr() + 4), INVALID),                                  # the image must be 16-byte aligned
                            (job(scale=4, flags=1 | 64), INVALID),            # 4x4 samples at scale 4 would need a sampling scale of 16
                            (job(n_frames=3), CAPACITY)):                     # more frames than the context was created for (scratch graphs)
            assert L.par_remaster_device(c.handle, C.byref(bad)) == status
            assert len(L.par_last_error(c.handle)) > 0
        assert L.par_remaster_device(c.handle, None) == INVALID
        assert L.par_border_walks(c.handle, None, None, 40, 30, 2, None, None, None, 10, None) == INVALID
        assert c.launch_count == launches                                     # nothing ran
        assert torch.equal(c.remaster(frames, scale=2, subdivide=True)["rgba"], good)


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [1, 2, 3, 4, 5, 6, 7, 8])
def test_output_formats_hold_the_same_image(lib, oracle, scale):
    """PAR_OUT_BGR8 (the 3-channel image Image::saveImage writes, Image.cpp:64-71) and PAR_OUT_INDEX8 (palette indices:
    every output pixel is a source colour, kernel.cu:98-101, or the background, main.cpp:260) are the RGBA8 image in
    another layout: byte for byte on the TMA and plain-load paths, on the exact path, flipped, with padded rows, through
    the stage entry and the host entry; the palette is ascending, starts with black and counts the frame's colours."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    frames_np = np.concatenate([synth.snes_stream(2, 72, 40, first_seed=77), synth.adversarial_sprite(72, 40, 3)[None]])
    F, H, W = frames_np.shape[:3]
    padded = np.zeros((F, H, 3 * W + 8), np.uint8)
    padded[:, :, :3 * W] = frames_np.reshape(F, H, 3 * W)
    padded[:, :, 3 * W:] = 201  # row padding is never a pixel
    with lib.Remaster(0, W, H, F) as c:
        for src in ("dense", "padded"):
            if src == "dense":
                frames = torch.from_numpy(frames_np).cuda()
            else:
                frames = torch.from_numpy(padded).cuda().as_strided((F, H, W, 3), (H * (3 * W + 8), 3 * W + 8, 3, 1))
            for flip in (False, True):
                for no_tma in (False, True):
                    ref = c.remaster(frames, scale=scale, subdivide=True, want=("rgba", "graph"), flip_output=flip, no_tma=no_tma)
                    rgba = ref["rgba"].cpu().numpy()
                    if scale in (4, 5) and not flip and not no_tma:
                        assert np.array_equal(rgba[0], oracle.pipeline(frames_np[0], scale=scale, want=("raster",))["raster"])
                    bgr = c.remaster(frames, scale=scale, subdivide=True, flip_output=flip, no_tma=no_tma, out_format=lib.OUT_BGR8)["rgba"]
                    assert bgr.shape == (F, scale * H, scale * W, 3)
                    assert np.array_equal(bgr.cpu().numpy(), rgba[..., [2, 1, 0]]), (src, flip, no_tma)
                    idx = c.remaster(frames, scale=scale, subdivide=True, flip_output=flip, no_tma=no_tma, out_format=lib.OUT_INDEX8)
                    assert idx["rgba"].shape == (F, scale * H, scale * W)
                    assert np.array_equal(lib.Remaster.expand_indexed(idx["rgba"], idx["palette"]), rgba), (src, flip, no_tma)
                    pal = idx["palette"].cpu().numpy().view(np.uint32)
                    cnt = idx["palette_count"].cpu().numpy()
                    for k in range(F):
                        cols = frames_np[k].reshape(-1, 3).astype(np.uint32)
                        words = np.unique(np.concatenate([[0], cols[:, 2] | cols[:, 1] << 8 | cols[:, 0] << 16]))
                        assert cnt[k] == len(words)
                        assert np.array_equal(pal[k, :cnt[k]] & 0xFFFFFF, words) and np.all(pal[k] >> 24 == 255)
                    # the stage entry and the exact slow path
                    st = c.raster(frames, ref["graph"], scale, True, flip_output=flip, no_tma=no_tma, out_format=lib.OUT_INDEX8)
                    assert torch.equal(st["rgba"], idx["rgba"]) and torch.equal(st["palette"], idx["palette"])
                    if not no_tma:
                        wide = c.raster(frames, ref["graph"], scale, True, flip_output=flip, debug_wide=True, out_format=lib.OUT_BGR8)
                        assert torch.equal(wide, bgr)
                        widx = c.raster(frames, ref["graph"], scale, True, flip_output=flip, debug_wide=True, out_format=lib.OUT_INDEX8)
                        assert torch.equal(widx["rgba"], idx["rgba"])
        host_in = torch.from_numpy(frames_np).pin_memory()
        rgba = c.remaster_host(host_in, scale=scale, subdivide=True)["rgba"].numpy()
        hb = c.remaster_host(host_in, scale=scale, subdivide=True, out_format=lib.OUT_BGR8)["rgba"].numpy()
        hi = c.remaster_host(host_in, scale=scale, subdivide=True, out_format=lib.OUT_INDEX8)
        assert np.array_equal(hb, rgba[..., [2, 1, 0]])
        assert np.array_equal(lib.Remaster.expand_indexed(hi["rgba"], hi["palette"]), rgba)
        # anti-aliased output in BGR8
        if scale <= 4:
            c.aa = 2
            a = c.remaster(torch.from_numpy(frames_np).cuda(), scale=scale, subdivide=True)["rgba"].cpu().numpy()
            b = c.remaster(torch.from_numpy(frames_np).cuda(), scale=scale, subdivide=True, out_format=lib.OUT_BGR8)["rgba"].cpu().numpy()
            assert np.array_equal(b, a[..., [2, 1, 0]])
            c.aa = 1


@pytest.mark.gpu
def test_indexed_output_of_frames_with_many_colours(lib):
    """A frame with exactly 256 colours (incl. black) is representable, one with more reports its count (> 256) and
    leaves the other frames of the batch intact; a frame larger than one palette chunk (64 K pixels) merges its chunks."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    rng = np.random.default_rng(5)
    H, W = 48, 64
    pal255 = np.unique(rng.integers(1, 1 << 24, 400, dtype=np.uint32))[:255]          # 255 colours + black = 256
    pal300 = np.unique(rng.integers(1, 1 << 24, 500, dtype=np.uint32))[:300]

    def frame(pal, h=H, w=W):
        idx = rng.integers(0, len(pal), (h, w))
        idx.reshape(-1)[:len(pal)] = np.arange(len(pal))   # every colour occurs
        c = pal[idx]
        return np.stack([(c >> 16) & 255, (c >> 8) & 255, c & 255], -1).astype(np.uint8)

    frames_np = np.stack([frame(pal255), frame(pal300), frame(pal255[:17])])
    with lib.Remaster(0, W, H, 3) as c:
        frames = torch.from_numpy(frames_np).cuda()
        rgba = c.remaster(frames, scale=4, subdivide=True)["rgba"].cpu().numpy()
        idx = c.remaster(frames, scale=4, subdivide=True, out_format=lib.OUT_INDEX8)
        cnt = idx["palette_count"].cpu().numpy()
        assert cnt[0] == 256 and cnt[1] > 256 and cnt[2] == 18, cnt
        full = lib.Remaster.expand_indexed(idx["rgba"], idx["palette"])
        assert np.array_equal(full[0], rgba[0]) and np.array_equal(full[2], rgba[2])
    big = frame(pal255[:40], 300, 400)                      # 120 000 pixels: two chunks
    with lib.Remaster(0, 400, 300, 1) as c:
        frames = torch.from_numpy(big[None]).cuda()
        rgba = c.remaster(frames, scale=2, subdivide=True)["rgba"].cpu().numpy()
        idx = c.remaster(frames, scale=2, subdivide=True, out_format=lib.OUT_INDEX8)
        assert int(idx["palette_count"][0]) == 41
        assert np.array_equal(lib.Remaster.expand_indexed(idx["rgba"], idx["palette"]), rgba)


@pytest.mark.gpu
def test_8x_output_at_every_16_byte_alignment(lib):
    """At 8x a lane's row segment is 32 bytes and leaves by one 256-bit store when the image base is 32-byte aligned,
    by two 128-bit stores otherwise (the C ABI asks for 16-byte alignment only): both give the same image."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    F, W, H, S = 3, 72, 40, 8
    frames = torch.from_numpy(synth.snes_stream(F, W, H, first_seed=4242)).cuda()
    n = F * S * H * S * W * 4
    with lib.Remaster(0, W, H, F) as c:
        want = c.remaster(frames, scale=S, subdivide=True)["rgba"]
        buf = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        base = buf.data_ptr()
        for off in ((-base) % 32, (-base) % 32 + 16):
            view = buf[off:off + n].view(F, S * H, S * W, 4)
            assert view.data_ptr() % 16 == 0 and (view.data_ptr() % 32 == 0) == (off == (-base) % 32)
            buf.zero_()
            c.remaster(frames, scale=S, subdivide=True, out={"rgba": view})
            assert torch.equal(view, want), off


@pytest.mark.gpu
def test_streams_devices_and_sub_batches(lib):
    """A context follows torch's current stream call by call (outputs allocated under `with torch.cuda.stream(s)` are
    produced on s), leaves the caller's current device alone, refuses tensors of another device, and gives the same
    result whether a batch runs as one launch per stage or in rounds of a few frames."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    frames_np = synth.snes_stream(7, 64, 48, first_seed=900)
    frames = torch.from_numpy(frames_np).cuda()
    with lib.Remaster(0, 64, 48, 7) as c:
        want = c.remaster(frames, scale=3, subdivide=True, want=("rgba", "graph", "labels"))
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        for _ in range(3):
            with torch.cuda.stream(s):
                # a long-running producer on s, then the remaster of ITS output on s: a context still bound to the
                # default stream would read the frames before the copy has happened
                staged = torch.zeros_like(frames)
                big = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
                for _k in range(4):
                    big.fill_(1)
                staged.copy_(frames)
                got = c.remaster(staged, scale=3, subdivide=True, want=("rgba", "graph", "labels"))
            s.synchronize()
            for k in want:
                assert torch.equal(got[k], want[k]), k
        for sub in (1, 2, 3, 7, 100):
            c.set_sub_batch(sub)
            got = c.remaster(frames, scale=3, subdivide=True, want=("rgba", "graph", "labels"))
            idx = c.remaster(frames, scale=3, subdivide=True, out_format=lib.OUT_INDEX8)
            for k in want:
                assert torch.equal(got[k], want[k]), (sub, k)
            assert np.array_equal(lib.Remaster.expand_indexed(idx["rgba"], idx["palette"]), want["rgba"].cpu().numpy())
        c.set_sub_batch(0)
        assert torch.cuda.current_device() == 0
        with pytest.raises(ValueError):
            c._job(_other_device(), 3, 0)


def _other_device():
    """A stand-in tensor that claims to live on another CUDA device (a one-GPU box cannot make a real one)."""
    class Fake:
        is_cuda = True
        device = __import__("torch").device("cuda", 1)
    return Fake()

"""The oracle against committed golden vectors (tests/golden/, generated FROM THE REFERENCE by
tests/golden/make_golden.py) and against the reference's own embedded fixtures: the two graph dumps
in kernel.cu:291-296 and the border walks of alex_png.txt.  Runs anywhere (no /root/reference)."""
import json
import os

import numpy as np
import pytest

from conftest import valid_vertex_mask

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FRAME_FILES = sorted(f for f in os.listdir(GOLD) if f.startswith("frames_"))


def _frame(z):
    H, W, ws = int(z["height"]), int(z["width"]), int(z["widthstep"])
    raw = np.ascontiguousarray(z["raw_rows"])
    return np.lib.stride_tricks.as_strided(raw, shape=(H, W, 3), strides=(ws, 3, 1)), raw


def test_yuv_words(oracle):
    z = np.load(os.path.join(GOLD, "yuv_words.npz"))
    for c, f, p in zip(z["colour"], z["fused"], z["plain"]):
        b = (int(c) & 255, (int(c) >> 8) & 255, int(c) >> 16)
        assert oracle.yuv_word(*b, True) == int(f)
        assert oracle.yuv_word(*b, False) == int(p)


def test_cells_4096(oracle):
    z = np.load(os.path.join(GOLD, "cells_4096.npz"))
    for key in range(4096):
        node = key & 255
        left = (4 if key & 256 else 0) | (128 if key & 512 else 0)
        right = (1 if key & 1024 else 0) | (32 if key & 2048 else 0)
        xy, n = oracle.cell_hull(node, left, right)
        assert n == int(z["count"][key])
        assert np.array_equal(xy * 4, z["verts_q4"][key, : n + 1].astype(np.float32)), key


@pytest.mark.parametrize("fname", FRAME_FILES)
def test_frame_fixture_every_stage(oracle, fname):
    z = np.load(os.path.join(GOLD, fname))
    img, _keep = _frame(z)
    for sub in (0, 1):
        got = oracle.pipeline(img, bool(sub), True, 4, ("graph_aux", "graph", "hull", "hull_count", "poly", "poly_count", "tri", "ntri"))
        assert np.array_equal(got["graph_aux"], z["graph_aux"])
        assert np.array_equal(got["graph"], z["graph"])
        assert np.array_equal(got["hull_count"], z["hull_count"])
        m = valid_vertex_mask(got["hull_count"], closing=True)[:, :9]
        assert np.array_equal((got["hull"][:, :9] * 4)[m], z["hull_q4"].astype(np.float32)[m])
        assert np.array_equal(got["poly_count"], z["poly_count_sub%d" % sub])
        m = valid_vertex_mask(got["poly_count"])[:, :16]
        assert np.array_equal((got["poly"][:, :16] * 64)[m], z["poly_q64_sub%d" % sub].astype(np.float32)[m])
        m = (np.arange(45)[None, :] < 3 * got["ntri"][:, None])[:, :42]
        assert np.array_equal((got["tri"][:, :42] * 64)[m], z["tri_q64_sub%d" % sub].astype(np.float32)[m])


def test_reference_graph_dumps_are_consistent(oracle):
    """kernel.cu:291-296: symmetric, no edges leaving the image, no crossing left."""
    z = np.load(os.path.join(GOLD, "graph_dumps.npz"))
    di = [-1, 0, 1, -1, 1, -1, 0, 1]
    dj = [1, 1, 1, 0, 0, -1, -1, -1]
    for name, interior in (("c_pattern", 8), ("alex", 178)):
        g = z[name]
        H, W = g.shape
        assert int((g == 90).sum()) == interior
        for j in range(H):
            for i in range(W):
                for e in range(8):
                    if g[j, i] >> e & 1:
                        ni, nj = i + di[e], j + dj[e]
                        assert 0 <= ni < W and 0 <= nj < H
                        assert g[nj, ni] >> (7 - e) & 1
        out, n_amb = oracle.resolve_crossings(g)
        assert n_amb == 0 and np.array_equal(out, g)
        assert np.array_equal(oracle.trivial_crossings(g), g)


def test_cc_labels_on_alex_fixture(oracle):
    """Known-answer test for connected components: the reference's 24x24 'alex' graph dump, its border
    walks in alex_png.txt and what its (dead-code) walker returns when run on the dump."""
    z = np.load(os.path.join(GOLD, "graph_dumps.npz"))
    lab = oracle.cc_labels(z["alex"]).reshape(-1)
    sizes = {int(k): int(v) for k, v in zip(*np.unique(lab, return_counts=True))}
    assert sizes == {0: 125, 5: 95, 17: 147, 30: 7, 36: 7, 80: 51, 148: 2, 159: 3, 172: 5, 183: 8, 225: 4, 294: 35, 301: 57,
                     319: 2, 342: 5, 346: 9, 351: 8, 370: 2, 375: 4}
    walker = json.load(open(os.path.join(GOLD, "border_walks_alex.json")))  # cc_functions.cu run on the dump
    assert sorted(w[0] for w in walker) == sorted(sizes)                     # a component is named by its first node
    for w in walker:
        assert len(set(lab[w].tolist())) == 1 and lab[w[0]] == w[0]
    txt = [[v for v in row if v >= 0] for row in z["alex_walks"].tolist()]    # alex_png.txt:1-21
    assert len(txt) == 21
    starts_at_label = 0
    for w in txt:
        assert len(set(lab[w].tolist())) == 1                                 # every listed walk stays inside one component
        starts_at_label += int(lab[w[0]] == w[0])
    assert starts_at_label == 19                                              # the two others (35, 83) are inner borders


def test_stages_d_to_raster_from_reference_dumps(oracle):
    """Stages D-G from a GIVEN graph (the dumps pin those stages, not A-C): cells of the dumps equal the
    4096-entry golden table; triangle and polygon rasters agree."""
    z = np.load(os.path.join(GOLD, "graph_dumps.npz"))
    cells = np.load(os.path.join(GOLD, "cells_4096.npz"))
    for name in ("c_pattern", "alex"):
        g = z[name]
        H, W = g.shape
        hull, cnt = oracle.cells(g)
        flat = g.reshape(-1)
        for n in range(H * W):
            left = int(flat[n - 1]) if n > 0 else 0
            right = int(flat[n + 1]) if n + 1 < H * W else 0
            key = int(flat[n]) | (256 if left & 4 else 0) | (512 if left & 128 else 0) | (1024 if right & 1 else 0) | (2048 if right & 32 else 0)
            assert cnt[n] == cells["count"][key]
            assert np.array_equal(hull[n, : cnt[n] + 1] * 4, cells["verts_q4"][key, : cnt[n] + 1].astype(np.float32))
        img = np.zeros((H, W, 3), np.uint8)
        img[..., 0] = (np.arange(H * W).reshape(H, W) * 37) & 255  # arbitrary colours; the graph is given
        tri, nt = oracle.triangulate(hull, cnt, W, H)
        assert (nt == cnt - 2).all()
        assert np.array_equal(oracle.raster_triangles(img, 4, tri, nt), oracle.raster_polygons(img, 4, hull, cnt))


def _first_walks(oracle, case):
    """The reference walker's output on one golden graph, reduced to the first walk of every component."""
    g = np.array(case["graph"], np.uint8).reshape(case["H"], case["W"])
    lab = oracle.cc_labels(g).reshape(-1)
    first = {}
    for w in case["walks"]:
        if lab[w[0]] == w[0] and w[0] not in first:
            first[w[0]] = w
    return g, lab, first


def test_border_walks_equal_the_reference_walker(oracle):
    """the oracle's border walk per component against what the reference's own (dead-code) walker,
    cc_functions.cu:348-503, produced on 10 small graphs + its 24x24 'alex' dump (tests/golden/make_border_walks.py,
    make_golden.py): every first walk of a component is identical, node for node; components the reference drops are dropped."""
    cases = json.load(open(os.path.join(GOLD, "border_walks_small.json")))
    assert len(cases) >= 8
    compared = 0
    for case in cases:
        g, lab, first = _first_walks(oracle, case)
        mine = oracle.border_walks(g)
        assert set(mine) == set(first), case["name"]
        for s0, w in first.items():
            assert mine[s0] == w, (case["name"], s0)
            compared += 1
    assert compared > 250
    z = np.load(os.path.join(GOLD, "graph_dumps.npz"))
    walker = json.load(open(os.path.join(GOLD, "border_walks_alex.json")))
    mine = oracle.border_walks(z["alex"])
    assert len(mine) == 19 and all(mine[w[0]] == w for w in walker)


def test_product_yuv_word_on_the_host(oracle):
    """The product's packed-YUV conversion (csrc/common.cuh, the same function the kernels inline, here through the C ABI's
    host entry par_yuv_word): every grey (the 256-bit rounding table), EVERY colour whose 299 b0 + 587 b1 + 114 b2 is a
    multiple of 1000 (the only ones the FMA chain decides), and a random sample, against the oracle's fused conversion."""
    import pixel_art_remaster_gpu_b200 as par
    for v in range(256):
        assert par.yuv_word(v, v, v) == oracle.yuv_word(v, v, v, True), v
    b = np.arange(256)
    T = (299 * b[:, None, None] + 587 * b[None, :, None] + 114 * b[None, None, :])
    hits = np.argwhere(T % 1000 == 0)
    assert 15000 < len(hits) < 20000
    for b0, b1, b2 in hits:  # every one of them
        assert par.yuv_word(int(b0), int(b1), int(b2)) == oracle.yuv_word(int(b0), int(b1), int(b2), True), (b0, b1, b2)
    rng = np.random.default_rng(3)
    for col in rng.integers(0, 1 << 24, 20000):
        b0, b1, b2 = int(col) & 255, (int(col) >> 8) & 255, int(col) >> 16
        assert par.yuv_word(b0, b1, b2) == oracle.yuv_word(b0, b1, b2, True)


def test_numpy_graph_builder_equals_the_oracle(oracle):
    """tests/np_graph.py (used by the GPU suite to build the expected graph of the 4096 x 4096 all-colours frame) against the
    oracle's stages A + B on small frames, on a slab of the all-colours frame and on degenerate shapes."""
    from np_graph import all_colours_frame, graph_aux_from_yuv
    from pixel_art_remaster_gpu_b200 import synth
    table = oracle.yuv_all(True)

    def words(img):
        c = img.astype(np.uint32)
        return table[c[..., 0] | c[..., 1] << 8 | c[..., 2] << 16]

    cases = [synth.snes_frame(64, 48, 3), synth.adversarial_sprite(72, 60), synth.snes_frame(1, 1, 4), synth.snes_frame(40, 1, 5),
             synth.snes_frame(1, 40, 6), np.ascontiguousarray(all_colours_frame()[2000:2040, 1000:1300])]
    for img in cases:
        want = oracle.trivial_crossings(oracle.similarity_graph(img))
        assert np.array_equal(graph_aux_from_yuv(words(img)), want), img.shape

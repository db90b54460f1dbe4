"""The oracle (oracle/remaster_oracle.c) against the REFERENCE'S OWN CODE compiled as host C++
(oracle/_ref/libref_host*.so, built from /root/reference by oracle/Makefile).  Skipped where that
build is absent (e.g. a box without /root/reference and without the prebuilt oracle/_ref)."""
import numpy as np
import pytest

from conftest import valid_vertex_mask
from pixel_art_remaster_gpu_b200 import synth


def test_yuv_words_all_colours(oracle, ref_host, ref_host_plain):
    """All 2^24 colours: fused-Y oracle == the reference compiled with FMA contraction (what nvcc does on
    the device); plain oracle == the reference compiled without; they differ on 2490."""
    fused, plain = oracle.yuv_all(True), oracle.yuv_all(False)
    assert np.array_equal(fused, ref_host.yuv_all())
    assert np.array_equal(plain, ref_host_plain.yuv_all())
    assert int((fused != plain).sum()) == 2490


FRAMES = [
    ("g1_96x80", lambda: synth.snes_frame(96, 80, synth.BASE_SEED + 1)),
    ("g1_128x96", lambda: synth.snes_frame(128, 96, synth.BASE_SEED + 2)),
    ("g1_256x224", lambda: synth.snes_frame(256, 224, synth.BASE_SEED)),
    ("g5_192x160", lambda: synth.adversarial_sprite(192, 160)),
    ("odd_50x37", lambda: synth.snes_frame(50, 37, 5)),
    ("padded_50x37_ws152", lambda: synth.pad_rows(synth.snes_frame(50, 37, 5), 152)),
    ("tiny_3x2", lambda: synth.snes_frame(3, 2, 8)),
    ("row_40x1", lambda: synth.snes_frame(40, 1, 10)),
    ("col_1x40", lambda: synth.snes_frame(1, 40, 11)),
]


@pytest.mark.parametrize("name,make", FRAMES, ids=[f[0] for f in FRAMES])
@pytest.mark.parametrize("subdivide", [False, True], ids=["hull", "subdivided"])
def test_every_stage_equals_reference(oracle, ref_host, name, make, subdivide):
    img = make()
    want = ("graph_aux", "graph", "hull", "hull_count", "poly", "poly_count", "tri")
    ref = ref_host.pipeline(img, subdivide, want)
    got = oracle.pipeline(img, subdivide, True, 4, want + ("ntri",))
    assert np.array_equal(got["graph_aux"], ref["graph_aux"])          # stages A + B, bit-exact
    assert np.array_equal(got["graph"], ref["graph"])                  # stage C, bit-exact
    assert np.array_equal(got["hull_count"], ref["hull_count"])        # stage D
    m = valid_vertex_mask(ref["hull_count"], closing=True)
    assert np.array_equal(got["hull"][m], ref["hull"][m])
    assert np.array_equal(got["poly_count"], ref["poly_count"])        # stage E
    m = valid_vertex_mask(ref["poly_count"])
    assert np.array_equal(got["poly"][m], ref["poly"][m])
    assert (got["ntri"] == np.maximum(got["poly_count"] - 2, 0)).all()  # ear clipping never gives up
    m = np.arange(45)[None, :] < 3 * got["ntri"][:, None]
    assert np.array_equal(got["tri"][m], ref["tri"][m])                # stage F


@pytest.mark.parametrize("w,h", [(4, 4), (7, 5), (65, 33), (129, 35)])
def test_dark_pixels_on_the_border_equal_reference(oracle, ref_host, w, h):
    """Black and near-black pixels on the image border (what the graph kernel's zero-filled staging must not link to):
    the oracle's graph equals the reference's own routines there too."""
    rng = np.random.default_rng(1000 * w + h)
    palette = np.array([[0, 0, 0], [1, 2, 1], [4, 4, 4], [3, 0, 5], [200, 30, 90], [0, 0, 0], [2, 2, 2], [250, 250, 250]], np.uint8)
    for k in range(3):
        img = palette[rng.integers(0, len(palette) if k else 4, (h, w))]
        if k == 2:
            img[:] = 0
        img = np.ascontiguousarray(img)
        ref = ref_host.pipeline(img, False, ("graph_aux", "graph"))
        got = oracle.pipeline(img, False, True, 4, ("graph_aux", "graph"))
        assert np.array_equal(got["graph_aux"], ref["graph_aux"])
        assert np.array_equal(got["graph"], ref["graph"])


def test_plain_host_arithmetic_variant(oracle, ref_host_plain):
    """The unfused-Y variant of the oracle equals the reference built with -ffp-contract=off."""
    img = synth.snes_frame(96, 80, 123)
    ref = ref_host_plain.pipeline(img, True, ("graph_aux", "graph"))
    assert np.array_equal(oracle.trivial_crossings(oracle.similarity_graph(img, fused=False)), ref["graph_aux"])


def test_all_cells_equal_reference(oracle, ref_host_plain):
    for key in range(4096):
        node = key & 255
        left = (4 if key & 256 else 0) | (128 if key & 512 else 0)
        right = (1 if key & 1024 else 0) | (32 if key & 2048 else 0)
        a, na = oracle.cell_hull(node, left, right)
        b, nb = ref_host_plain.cell(node, left, right)
        assert na == nb and np.array_equal(a, b), key


def test_raster_rules_agree(oracle):
    """The two statements of the raster rule — paint the reference's triangle list with the top-left
    rule vs. even-odd test of the displaced sample against the polygon — give the same image."""
    for img, s in ((synth.snes_frame(96, 80, 31), 4), (synth.snes_frame(64, 48, 32), 8), (synth.adversarial_sprite(96, 80), 3)):
        r = oracle.pipeline(img, True, True, s, ("poly", "poly_count", "raster"))
        assert oracle.all_dyadic64(r["poly"])
        assert np.array_equal(oracle.raster_polygons(img, s, r["poly"], r["poly_count"]), r["raster"])
